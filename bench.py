#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: 1080p frames/s of the P-picture ME + transform/quant/reconstruction path.

Workload at N GPUs (config 5 of BASELINE.json, sharded as SURVEY.md §8e "batch of independent sequences"): every
GPU codes `--seqs` independent synthetic 1080p sequences (coded 1920x1072, IPPP, QP 28, WindowSize 32 = +-16 search,
MAXDIFF 3) in lockstep; one STEP = one P picture of every sequence on the GPU: scene-change SAD (a12), phases A/B/C
(a3-a10), dpb swap (a11), phase R (a2). Weak scaling: per-GPU work fixed, no data-path collective.

One step = fh264_upload_source_batch + fh264_encode_p_stream (include/fh264_b200.h): no host round trip per picture; the
scene-change IDR rule is decided on the device.

  value  : pictures/s with the source pictures already resident in HBM (device-to-device into `frame`)
  e2e    : the same through the C ABI with HOST buffers: pinned H2D of every source picture and, inside the timed region, D2H
           of what the reference's host code needs to write the slice NAL unit (entropy-coded slice data, 32 B/MB side
           information, status words); `device_cavlc.e2e_records`: the 832-byte macroblock records instead
  --impl reference : the unmodified reference encoder (oracle/_ref/ref_encoder, compiled from /root/reference in the
           build container) on the box's host cores, one single-threaded process per sequence.

The oracle / reference binary is used here ONLY as the timed CPU baseline, never by the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080            # input size; coded 1920x1072 after the reference's crop (fileIO.cpp:242-243)
QP, WINDOW, MAXDIFF = 28, 32, 3
CLIP_LEN = 6                          # distinct pictures per sequence (set per run: warm-up + steps + 3, at most CLIP_MAX); past the end the clip plays back
CLIP_MAX = 40
ALG_BYTES_PER_MB = 1984               # 384 src + 384 ref + 384 recon + 768 levels + 64 metadata
ALG_INTOPS_PER_MB = 0.43e6
LAUNCHES_PER_STEP = 14                # own kernels per step and group: source swap, begin, scene SAD, scene gate, stage 3, stage 2, phase S 2, B, C,
                                      # dpb swap, phase R 3 (the e2e step adds the 4 entropy-coding kernels; --pipeline: two halves of 15)
# The same 0.43 M lane-ops per macroblock split over the kernels that do them (SURVEY.md §8d's per-stage counts; 4 partitions
# per macroblock, 40 ops per feature-cost evaluation, 32 packed-byte ops per 8x8 SAD):
#   stage 3: 1,475 evaluations + 33 SADs; stage 1 (phase S): 400 evaluations + 17 SADs; stage 2: ~100 evaluations + 32 SADs;
#   phase C: two motion compensations + 48 4x4 transforms with quant/dequant; phase R: 16-plane interpolation + box sums.
ALG_INTOPS_BY_KERNEL = {"k_stage3": 4 * (1475 * 40 + 33 * 32), "k_spec": 4 * (400 * 40 + 17 * 32), "k_stage2": 4 * (100 * 40 + 32 * 32),
                        "k_phase_b": 1000, "k_phase_c": 24000, "k_interp": 110 * 256, "k_features": 190 * 256, "k_tile_index": 0}


def ncu_traffic(dom, seqs_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by profiles/tools/ncu_traffic.py); null when no capture matches this launch size."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, None, None
    if t.get("seqs_per_launch") != seqs_per_launch:
        return None, None, "profiles/ncu_traffic.json holds a capture at %s sequences per launch, this run uses %d" % (t.get("seqs_per_launch"), seqs_per_launch)
    by = {}
    for name, v in t["kernels"].items():
        short = name.split("<")[0]
        if short == "k_stage2" and "4096" in name:
            continue                                    # the (usually empty) fallback launch
        by[short] = v["traffic_bytes"]
    return by.get(dom), by, "profiles/ncu_traffic.json (%s)" % os.path.basename(t.get("source", "?"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def measured_int_peak(device):
    """Sustained integer-pipe rate of THIS GPU, measured now by the library's micro-benchmark (csrc/intpeak.cuh: 2048 threads
    per SM, 8 independent chains, 128 statements per iteration): the instructions the ME kernels are made of (VABSDIFF4.U8.ACC,
    VIADDMNMX[.S16x2], IMAD) issue at 64 lanes per SM per clock; only plain adds have the second 64 lanes. Returns
    (T lane-ops/s of that 64-lane pipe, the per-instruction figures)."""
    from h264_fer_b200 import native
    d = native.measure_int_peak(device)
    return float(min(d["imad"], d["viaddmnmx"], d["viaddmnmx_s16x2"])), d


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def set_clip_len(args):
    global CLIP_LEN
    CLIP_LEN = max(6, min(CLIP_MAX, max(args.warmup, 3) + args.steps + 3))


def pingpong(t, n):
    """0,1,..,n-1,n-2,..,1,0,1,.. : consecutive pictures always differ by exactly one pan step. (The clip is made long enough
    for warm-up + steps, so a default run only ever plays forward.)"""
    period = 2 * (n - 1)
    k = t % period
    return k if k < n else period - k


def make_clips(nseq, first_seed, length=None):
    from concurrent.futures import ThreadPoolExecutor
    from h264_fer_b200 import synth
    n = CLIP_LEN if length is None else length

    def one(i):
        c = synth.SynthClip(WIDTH, HEIGHT, first_seed + i)
        frames = []
        for t in range(n):
            y, cb, cr = c.frame(t)
            frames.append((synth.crop16(y), synth.crop16(cb, chroma=True), synth.crop16(cr, chroma=True)))
        return frames

    with ThreadPoolExecutor(max_workers=min(8, max(1, nseq))) as ex:
        return list(ex.map(one, range(nseq)))


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference_processes(nproc, pictures, first_seed, tmpdir):
    """nproc single-threaded reference encoders in parallel, each coding 1 I + (pictures-1) P pictures of its own 1080p
    clip. Returns per-process summaries (hot-path seconds per picture)."""
    from h264_fer_b200 import synth
    from oracle import refdump
    procs = []
    for i in range(nproc):
        y4m = os.path.join(tmpdir, "seq%d.y4m" % i)
        synth.write_y4m(y4m, WIDTH, HEIGHT, first_seed + i, pictures)
        cmd = [refdump.REF_ENCODER, y4m, os.path.join(tmpdir, "seq%d.264" % i), "-", str(pictures), str(QP), "0", str(WINDOW),
               str(MAXDIFF), "1000", "0"]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, cwd=tmpdir))
    outs = []
    for p in procs:
        out, _ = p.communicate()
        line = [l for l in out.decode().splitlines() if l.startswith("{")][-1]
        outs.append(json.loads(line))
    return outs


def hot_seconds(summary, pic):
    return summary["t_select"][pic] + summary["t_inter"][pic] + summary["t_tq"][pic] + summary["t_fill"][pic]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refdump
    cores = os.cpu_count() or 1
    nseq_total = args.seqs * args.gpus
    nproc = max(1, min(cores, nseq_total))
    steps = max(1, min(args.steps, 4))
    warm = 1 if args.warmup > 0 else 0
    line = {"impl": "reference", "metric": "1080p P-picture frames/s, ME + transform/quant/reconstruction hot path", "unit": "frames/s",
            "n_gpus": args.gpus, "higher_is_better": True, "scaling": "weak", "dtype": "u8/int32", "data": "synthetic", "vs_baseline": None,
            "config": workload_config(args, nseq_total)}
    if not refdump.have_ref_encoder():
        # the port (oracle/fh264_oracle.c) is the fallback CPU implementation of the path
        from oracle import port
        clips = make_clips(1, 100)
        o = port.Oracle(WIDTH, clips[0][0][0].shape[0])
        ref = clips[0][0]
        ts = []
        for t in range(1, warm + steps + 1):
            t0 = time.perf_counter()
            o.phase_r(ref[0])
            _, ref = o.encode_p(clips[0][pingpong(t, CLIP_LEN)], ref, QP, WINDOW, MAXDIFF)
            ts.append(time.perf_counter() - t0)
        sec = sum(ts[warm:])
        value = steps / sec
        line.update(value=value, steps=steps, warmup=warm, ms_per_step=1000 * sec / steps,
                    cpu_baseline={"value": value, "unit": "frames/s", "cores": 1, "kind": "port",
                                  "sample": "%d 1080p P pictures of 1 sequence, oracle port, 1 thread" % steps},
                    e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return 0
    tmp = tempfile.mkdtemp(prefix="fh264_refarm_")
    outs = run_reference_processes(nproc, 1 + warm + steps, 100, tmp)
    # one step = one P picture of every running process; hot-path time only, slowest process per step
    per_step = [max(hot_seconds(o, 1 + warm + k) for o in outs) for k in range(steps)]
    sec = sum(per_step)
    value = nproc * steps / sec
    line.update(value=value, steps=steps, warmup=warm, ms_per_step=1000 * sec / steps,
                cpu_baseline={"value": value, "unit": "frames/s", "cores": nproc, "kind": "reference",
                              "sample": "%d concurrent single-threaded reference processes (of %d host cores), each 1 I + %d P 1080p pictures; "
                                        "timed: %d P pictures per process, hot-path calls only (interEncoding, quantizationTransform, "
                                        "transformDecodingP_Skip, FillInterpolatedRefFrame, selectNALUnitType)" % (nproc, cores, warm + steps, steps)},
                e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))
    return 0


def workload_config(args, nseq_total):
    return {"workload": "BASELINE.json config 5 sharding: %d independent synthetic 1080p sequences per GPU (%d total), coded 1920x1072, "
                        "IPPP, 1 step = 1 P picture of every sequence" % (args.seqs, nseq_total),
            "qp": QP, "window": WINDOW, "maxdiff_set": MAXDIFF, "basic": 0, "seqs_per_gpu": args.seqs, "groups_per_gpu": getattr(args, "groups", 1),
            "clip": "%d distinct pictures per sequence played forward (pan 2x1 px per picture, moving square, noise); every timed region starts at picture 1" % CLIP_LEN,
            "first_picture": "source picture 0 uploaded as the reconstruction (I pictures are host work, out of scope)",
            "l2": "per-step working set (>330 MB of reference planes/features per sequence) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=8, help="independent sequences per GPU")
    ap.add_argument("--groups", type=int, default=1, help="sequence groups per GPU (session + stream + host thread each)")
    ap.add_argument("--threaded", action="store_true", help="drive even a single group from a worker thread")
    ap.add_argument("--pipeline", action="store_true", help="software-pipeline the two halves of every step (fh264_set_pipeline; measured: no gain, profiles/r02_pipeline.md)")
    ap.add_argument("--mode", default="sequences", choices=["sequences", "bands"],
                    help="sequences: independent sequences per GPU (weak scaling, the headline); bands: ONE 1080p sequence split into MB-row bands over all GPUs (BASELINE config 4, strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cavlc", action="store_true", help="skip the device-CAVLC line item (SURVEY.md §8(f) rank 1)")
    ap.add_argument("--no-intra", action="store_true", help="skip the device I-picture line item (SURVEY.md §8(f) rank 2)")
    ap.add_argument("--no-bands", action="store_true", help="at N > 1: skip the MB-row band measurement that follows the replica measurement")
    args = ap.parse_args()
    set_clip_len(args)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200 import sharding
    from h264_fer_b200.native import PinnedArray, StreamOut, ST_GATED_TOTAL

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))       # (NCCL may print its banner; the JSON line is the LAST line)
    if args.mode == "bands":
        return band_mode(args, rank, world, local)
    B, K, Wu = args.seqs, args.steps, max(args.warmup, 3)
    G = max(1, min(args.groups, B))                      # sequence groups: one session + stream + host thread each
    my_seqs = sharding.sequences_for_rank(B * world, rank, world)       # global sequence ids of this GPU (seed = 100 + id)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=8) as ex:
        clips = list(ex.map(lambda sid: make_clips(1, 100 + sid)[0], my_seqs))
    H = clips[0][0][0].shape[0]
    nmb = (WIDTH // 16) * (H // 16)
    ysz, csz = WIDTH * H, WIDTH * H // 4
    master = torch.cuda.current_stream()

    SLICE_COPY = 65536        # bytes of slice data per sequence and step that come home (a 1080p P slice at QP 28 is ~21 KB)
    pic = ysz + 2 * csz

    class Group:
        """A slice of the GPU's sequences driven through one session (the reference is one sequence per process; here a group
        advances its sequences in lockstep). One step = fh264_upload_source_batch + fh264_encode_p_stream: no host round trip."""

        def __init__(self, seq_ids):
            self.ids = seq_ids
            self.n = len(seq_ids)
            self.pinned = [PinnedArray((self.n, pic), np.uint8) for _ in range(CLIP_LEN)]      # picture t of every sequence: one block
            for t in range(CLIP_LEN):
                a = self.pinned[t].array
                for j, b in enumerate(seq_ids):
                    a[j, :ysz] = clips[b][t][0].ravel(); a[j, ysz:ysz + csz] = clips[b][t][1].ravel(); a[j, ysz + csz:] = clips[b][t][2].ravel()
            self.dev = [torch.from_numpy(self.pinned[t].array.copy()).cuda() for t in range(CLIP_LEN)]
            self.results = PinnedArray((self.n, nmb), fh.MB_RESULT_DTYPE)
            self.s = fh.Session(WIDTH, H, batch=self.n, device=local)
            self.stream = torch.cuda.Stream()
            self.s.set_stream(self.stream.cuda_stream)
            # what a step sends home: "dev" the status words only; "host" the entropy-coded slice data + 32 B/MB side information
            # (what the reference's host code needs to write the NAL unit); "records" the 832-byte macroblock records instead
            self.outs = {"dev": StreamOut(self.n, nmb), "host": StreamOut(self.n, nmb, slice_bytes=SLICE_COPY, mb_info=True),
                         "records": StreamOut(self.n, nmb, records=True)}
            self.t = 1

        def reset(self, mode="dev"):
            self.s.set_pipeline(1 if args.pipeline else 0)
            for j, b in enumerate(self.ids):
                self.s.upload_recon(j, *clips[b][0])
            self.s.sync()
            self.t = 1
            self.upload(mode)                                    # prime: the first picture to code
            self.s.sync()

        def upload(self, mode):
            """Hands the next picture of every sequence to the library in one call (pinned H2D, or D2D for "dev"): the copies run on
            the library's upload stream into the source buffers that are not being coded."""
            k = pingpong(self.t, CLIP_LEN)
            self.t += 1
            if mode == "dev":
                self.s.upload_source_batch(self.dev[k].data_ptr(), pic, device=True)
            else:
                self.s.upload_source_batch(self.pinned[k].ptr, pic)

        def step(self, mode):
            """One step = code the picture uploaded last — IDR decision on the device (selectNALUnitType's scene-change rule,
            ref_frames.cpp:210-224), phases A / B / C, entropy coding when the slice data goes home, dpb swap, phase R — then
            upload the next one (its H2D overlaps this picture's coding)."""
            self.s.encode_p_stream(QP, WINDOW, MAXDIFF, 0, scene_gate=True, out=self.outs[mode])
            self.upload(mode)

        def gated_total(self):
            st = self.outs["dev"].status.array
            return int(sum(int(st[j, ST_GATED_TOTAL]) for j in range(self.n)))

    groups = [Group(list(range(B))[i::G]) for i in range(G)]

    def timed(mode):
        host = mode != "dev"
        for gr in groups:
            gr.reset(mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event() for _ in groups]
        bar_warm, bar_go = threading.Barrier(G + 1), threading.Barrier(G + 1)
        errors = []

        def worker(gr, ev):
            try:
                torch.cuda.set_device(local)
                for _ in range(Wu):
                    gr.step(mode)
                gr.s.sync()
                bar_warm.wait()
                bar_go.wait()
                gr.stream.wait_event(e0)                         # the timed region starts at e0 on every group stream
                for _ in range(K):
                    gr.step(mode)
                gr.s.sync()                                      # everything of the last step is done / home inside the timed region
                ev.record(gr.stream)
            except Exception as ex:       # surfaced by the main thread
                errors.append(ex)
                try:
                    bar_warm.abort(); bar_go.abort()
                except Exception:
                    pass

        inline = G == 1 and not args.threaded            # one group: drive it from the main thread
        threads = [] if inline else [threading.Thread(target=worker, args=(gr, ev)) for gr, ev in zip(groups, ends)]
        for th in threads:
            th.start()
        if inline:
            for _ in range(Wu):
                groups[0].step(mode)
            groups[0].s.sync()
        else:
            bar_warm.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0.record(master)
        if inline:
            groups[0].stream.wait_event(e0)
            for _ in range(K):
                groups[0].step(mode)
            groups[0].s.sync()                                   # the pipeline lanes and the last copies home are inside the timed region
            ends[0].record(groups[0].stream)
        else:
            bar_go.wait()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        for ev in ends:
            master.wait_event(ev)
        e1.record(master)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = sharding.reduce_max(e0.elapsed_time(e1))          # max over ranks
        for gr in groups:
            for j in range(gr.n):
                gr.s.picture_status(j)
        return ms, clocks

    ms_dev, clocks = timed("dev")
    idr_decisions = sum(sum(v) for v in sharding.gather_counts([gr.gated_total() for gr in groups]))      # pictures the device-side scene gate stopped, all ranks
    ms_e2e, clocks_e2e = timed("host")
    g0 = groups[0]
    slice_bits = [int(g0.outs["host"].slice_stat.array[j, 1]) for j in range(g0.n)]
    slice_flags = [int(g0.outs["host"].slice_stat.array[j, 0]) for j in range(g0.n)]
    if any(slice_flags) or max(slice_bits) > 8 * SLICE_COPY or min(slice_bits) <= 0:
        raise SystemExit("device CAVLC: flags %s, bits %s (copy bound %d bytes)" % (slice_flags, slice_bits, SLICE_COPY))
    # the record path (832 B per macroblock home instead of the slice data): line item beside the headline
    ms_rec, _ = timed("records") if not args.no_cavlc else (None, None)
    cavlc = {"slice_bytes_last_step_group0": int(sum((b + 7) // 8 for b in slice_bits)),
             "note": "e2e returns what the reference's host code needs to write the slice NAL unit: the entropy-coded slice data (device CAVLC, "
                     "cavlc.cuh; %d bytes per sequence copied home) + 32 B/MB side information + the status words; e2e_records returns the 832-byte "
                     "macroblock records instead (host-side entropy coding)" % SLICE_COPY}
    if ms_rec is not None:
        cavlc["e2e_records"] = {"value": B * world * K / (ms_rec / 1000.0), "unit": "frames/s", "ms_per_step": ms_rec / K, "d2h_bytes_per_step": B * nmb * 832 + B * 96}

    # device I-picture line item: every sequence of group 0 codes its current source picture as an IDR
    # picture on the device (fh264_encode_i: intra mode searches, both CAVLC bit-cost trials, TQ, reconstruction, then phase R),
    # pinned H2D of the pictures and D2H of the 832-byte records inside the timed region. Reported beside the headline, not in it.
    intra = None
    if not args.no_intra and G == 1:
        g0 = groups[0]
        i_out = PinnedArray((g0.n, nmb), fh.MB_RESULT_I_DTYPE)
        ker_ms, call_ms = [], []
        for it in range(1 + 3):
            g0.t = 1 + it
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g0.upload("host")
            g0.s._ck(g0.s.L.fh264_encode_i(g0.s.handle, 0, g0.n, QP, i_out.ptr))
            if it:
                call_ms.append(1000.0 * (time.perf_counter() - t0))
                ker_ms.append(g0.s.last_intra_ms())
        km, cm = sum(ker_ms) / len(ker_ms), sum(call_ms) / len(call_ms)
        intra = {"e2e_value": g0.n / (cm / 1000.0), "unit": "I pictures/s (this GPU)", "call_ms": cm, "k_intra_ms": km, "pictures_per_launch": g0.n,
                 "intra4x4_share": float((i_out.array["mb_type"] == 0).mean()), "gpu_launches_per_call": 7,
                 "note": "fh264_encode_i on the group's pictures: k_intra (wavefront, one warp per macroblock) + dpb swap + phase R; call_ms is host wall "
                         "clock around upload + the synchronous call (H2D of the pictures, D2H of the records included), k_intra_ms the CUDA-event time of "
                         "the wavefront kernel; the reference codes one 1080p I picture in about 1.7-3.3 s on one host core"}

    # per-kernel device times (CUDA events inside the library, on the launching stream): group 0 alone, a few steps
    g0 = groups[0]
    g0.reset()
    g0.s.set_pipeline(0)              # the per-kernel split is measured unpipelined: every kernel alone on the GPU, all sequences per launch
    acc, nsamp = {}, 0
    for i in range(3 + 4):
        g0.s.encode_p(QP, WINDOW, MAXDIFF, 0, sync=False, download=False)
        g0.upload("dev")
        tm = g0.s.last_timings()
        if i >= 3:
            for k_, v in tm.items():
                acc[k_] = acc.get(k_, 0.0) + v
            nsamp += 1
    tm = {k_: v / nsamp for k_, v in acc.items()}
    counts = [g0.s.mode_counts(0)]
    Bk = g0.n                                             # pictures per kernel launch in the instrumented run

    # BASELINE config 4 at N > 1: ONE 1080p sequence split into MB-row bands over the GPUs of the node (strong scaling), measured
    # after the replica measurement so that the driver's scaling record carries both partitions the north star names.
    for gr in groups:
        gr.s.close()
    groups = []
    band = None
    if not args.no_bands:
        try:
            band = band_measure(args, rank, world, local)      # N = 1: the single-sequence latency the band split is measured against
        except Exception as ex:                                # (the replica measurement above stands on its own)
            band = {"error": repr(ex)}

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        frames_per_step = B * world
        value = frames_per_step * K / (ms_dev / 1000.0)
        e2e_value = frames_per_step * K / (ms_e2e / 1000.0)
        kernels = {"k_stage3": tm["k_stage3_ms"], "k_stage2": tm["k_stage2_ms"], "k_spec": tm["k_spec_ms"], "k_phase_b": tm["phase_b_ms"],
                   "k_phase_c": tm["phase_c_ms"], "k_interp": tm["k_interp_ms"], "k_features": tm["k_features_ms"], "k_tile_index": tm["k_tile_index_ms"]}
        dom = max(kernels, key=kernels.get)
        alg_bytes = ALG_BYTES_PER_MB * nmb * Bk           # one launch processes the group's pictures
        achieved = alg_bytes / (kernels[dom] / 1000.0) / 1e9
        ipk, ipk_detail = measured_int_peak(local)        # T lane-ops/s of the 64-lane integer pipe, measured on this GPU in this run
        dom_ops = ALG_INTOPS_BY_KERNEL[dom] * nmb * Bk    # algorithmic lane-ops of one launch of the dominant kernel
        dom_tops = dom_ops / (kernels[dom] / 1000.0) / 1e12
        # phase C alone is the HBM-bound kernel of the path: report it too
        c_achieved = alg_bytes / (tm["phase_c_ms"] / 1000.0) / 1e9
        int_ops = ALG_INTOPS_PER_MB * nmb * B * world * K          # whole timed job
        step_ms = ms_dev
        traffic, traffic_by_kernel, traffic_src = ncu_traffic(dom, Bk)
        line = {
            "metric": "1080p P-picture frames/s, ME + transform/quant/reconstruction hot path", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": Wu, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic", "config": workload_config(args, B * world),
            "macroblocks_per_s": value * nmb,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * (ysz + 2 * csz), "d2h_bytes_per_step": B * (SLICE_COPY + nmb * 32 + 8 + 96),
                    "ms_per_step": ms_e2e / K,
                    "returns": "device-CAVLC slice data + 32 B/MB side information + status words (see device_cavlc)"},
            "gpu_launches": LAUNCHES_PER_STEP * K * G,
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": {"bound": "int", "kernel": dom, "achieved": dom_tops, "peak": ipk, "unit": "T int-lane-op/s", "frac": dom_tops / ipk,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "measured in this run on this GPU (fh264_measure_int_peak: IMAD / VIADDMNMX / VIADDMNMX.S16x2 issue rate = 64 lanes per SM per clock)",
                         "peak_detail_tops": ipk_detail,
                         "note": "the motion-search kernels are bound by the integer pipe, not by HBM (SURVEY.md 8d): achieved = algorithmic lane-ops of the dominant "
                                 "kernel (%d per MB, the reference's own evaluation counts) x %d MB x %d pictures per launch / its live CUDA-event duration; "
                                 "whole step: int_roofline; every kernel: int_roofline_by_kernel; the HBM-shaped kernel: roofline_phase_c; HBM fraction of "
                                 "the dominant kernel: roofline_hbm"
                                 % (ALG_INTOPS_BY_KERNEL[dom], nmb, Bk)
                                 + ("; k_stage2's algorithmic count is the ~100 candidate evaluations + 32 SADs per partition only - SURVEY.md 8d deliberately does "
                                    "not count the SEARCH for the candidates (the reference scans ~83 k bucket entries per partition, this kernel ~2.1 k index "
                                    "entries), which is what the kernel spends its time on" if dom == "k_stage2" else "")},
            "roofline_hbm": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src},
            "roofline_phase_c": {"bound": "hbm", "kernel": "k_phase_c", "achieved": c_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": c_achieved / hbm_peak},
            "int_roofline_by_kernel": {k_: {"ops_per_mb": ALG_INTOPS_BY_KERNEL[k_], "ms": kernels[k_],
                                            "frac": (ALG_INTOPS_BY_KERNEL[k_] * nmb * Bk / (kernels[k_] / 1000.0) / 1e12) / ipk if kernels[k_] > 0 else None}
                                       for k_ in kernels},
            "int_roofline": {"ops_per_mb": ALG_INTOPS_PER_MB, "achieved_tops": int_ops / (step_ms / 1000.0) / 1e12 / world,
                             "peak_tops_nominal": 148 * 128 * sm_max * 1e6 / 1e12,
                             "frac": (int_ops / (step_ms / 1000.0) / 1e12 / world) / (148 * 128 * sm_max * 1e6 / 1e12),
                             "note": "SURVEY.md §8d: 0.43 M int32-lane ops per MB over the whole timed job, per GPU; peak = 148 SMs x 128 lanes x max SM clock; "
                                     "peak_tops_measured = sustained rate of the 64-lane integer pipe the ME instructions issue on, measured in this run"},
            "dram_traffic_bytes_per_launch": traffic_by_kernel,
            "kernel_ms_per_step": kernels, "phase_ms_per_step": {k_: tm[k_] for k_ in ("phase_a_ms", "phase_b_ms", "phase_c_ms", "copy_phase_r_ms", "total_ms")},
            "idr_decisions": idr_decisions, "mode_counts_last_picture_seq0": counts[0],
        }
        line["int_roofline"]["peak_tops_measured"] = ipk
        line["int_roofline"]["frac_measured"] = line["int_roofline"]["achieved_tops"] / ipk
        if band is not None:
            line["band_mode"] = band
        if cavlc is not None:
            line["device_cavlc"] = cavlc
        if intra is not None:
            line["device_intra"] = intra
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        sys.stdout.flush()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def band_mode(args, rank, world, local):
    """--mode bands: only the band measurement, as its own JSON line."""
    import torch.distributed as dist
    r = band_measure(args, rank, world, local)
    if rank == 0:
        print(json.dumps({"metric": "1080p P-picture frames/s, ME + transform/quant/reconstruction hot path (ONE sequence, MB-row bands)",
                          "value": r["frames_s"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_picture"],
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
                          "config": {"workload": r["workload"], "qp": QP, "window": WINDOW, "maxdiff_set": MAXDIFF}, "rank0_phase_ms": r["rank0_phase_ms"]}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def band_measure(args, rank, world, local):
    """BASELINE config 4: one 1080p sequence, every P picture split into MB-row bands over the GPUs of the node. Returns a dict
    (same on every rank)."""
    import torch
    import torch.distributed as dist
    import h264_fer_b200 as fh
    from h264_fer_b200 import sharding
    from h264_fer_b200.bands import BandSession
    from h264_fer_b200.native import PinnedArray
    K, Wu = args.steps, max(args.warmup, 3)
    clip = make_clips(1, 100)[0]
    H = clip[0][0].shape[0]
    nmb = (WIDTH // 16) * (H // 16)
    ysz, csz = WIDTH * H, WIDTH * H // 4
    dev = [torch.from_numpy(np.concatenate([p.ravel() for p in clip[t]])).cuda() for t in range(CLIP_LEN)]
    results = PinnedArray((1, nmb), fh.MB_RESULT_DTYPE)
    bs = BandSession(WIDTH, H, device=local, rank=rank, world=world)
    stream = torch.cuda.Stream()
    bs.s.set_stream(stream.cuda_stream)
    bs.upload_recon(*clip[0])
    t = 1

    from h264_fer_b200.native import StreamOut
    sout = StreamOut(1, nmb)

    def step():
        # one picture: scene SAD measured on the device (every rank holds the whole picture: same value everywhere), phases A / B / C
        # on this rank's band, reconstruction exchange, picture barrier, phase R — no host round trip
        nonlocal t
        p = dev[pingpong(t, CLIP_LEN)].data_ptr(); t += 1
        bs.s.upload_source_batch(p, ysz + 2 * csz, device=True)
        bs.s.encode_p_stream(QP, WINDOW, MAXDIFF, 0, scene_gate=2, out=sout)

    for _ in range(Wu):
        step()
    bs.s.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = sharding.reduce_max(e0.elapsed_time(e1))
    bs.s.picture_status(0)
    tm = bs.s.last_timings()
    out = {"ms_per_picture": ms / K, "frames_s": K / (ms / 1000.0), "bands_mb_rows": [b - a for a, b in bs.bands], "rank0_phase_ms": tm,
           "workload": "BASELINE.json config 4: one synthetic 1080p sequence, MB-row bands %s over %d GPUs, NVLink peer-memory wavefront hand-off and "
                       "reconstruction exchange fused into the kernels (no NCCL call on the data path)" % ([b - a for a, b in bs.bands], world)}
    bs.close()
    return out


def cpu_baseline():
    """Reference (or port) on ONE host core, bounded sample: 1 I + 1 P 1080p picture; hot-path time of the P picture."""
    from oracle import refdump
    try:
        if refdump.have_ref_encoder():
            tmp = tempfile.mkdtemp(prefix="fh264_cpubase_")
            o = run_reference_processes(1, 2, 100, tmp)[0]
            sec = hot_seconds(o, 1)
            return {"value": 1.0 / sec, "unit": "frames/s", "cores": 1, "kind": "reference",
                    "sample": "unmodified reference, 1 thread: 1 I + 1 P 1080p picture (seed 100), timed = hot-path calls of the P picture "
                              "(%.2f s; whole picture %.2f s); 1 sample" % (sec, o["t_picture"][1])}
        from oracle import port
        clips = make_clips(1, 100)
        o = port.Oracle(WIDTH, clips[0][0][0].shape[0])
        t0 = time.perf_counter()
        o.phase_r(clips[0][0][0])
        o.encode_p(clips[0][1], clips[0][0], QP, WINDOW, MAXDIFF)
        sec = time.perf_counter() - t0
        return {"value": 1.0 / sec, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "oracle port, 1 thread, 1 P 1080p picture (%.2f s)" % sec}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}


if __name__ == "__main__":
    sys.exit(main())
