/* Build shim: the reference declares OpenCL handle types in openCL_functions.h and ref_frames.cpp.
 * The oracle build keeps OpenCL disabled (OpenCLEnabled=false => the reference's CPU intra path),
 * so opaque pointer types are all that is needed. Test infrastructure only. */
#pragma once
typedef void *cl_mem;
typedef void *cl_command_queue;
typedef void *cl_context;
typedef void *cl_kernel;
