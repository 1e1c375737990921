// Oracle build: OpenCL lifecycle seam of the reference (openCL_functions.h:4-19) stubbed OFF so that the
// reference runs its own CPU path everywhere (the GPU intra path produces different bitstreams,
// intra.cpp:961-977 vs :978-1049). Test infrastructure only.
#include <CL/cl.h>
bool OpenCLEnabled = false;
int *predModes16x16 = 0, *predModes4x4 = 0;
cl_mem frame_mem, dpb_mem, ans_mem;
cl_command_queue cmd_queue;
cl_context context;
cl_kernel kernel[2];
void InitCL() {}
void CloseCL() {}
void AllocateFrameBuffersCL() {}
void IntraCL() {}
void WaitIntraCL(int) {}
void subtractFramesCL(unsigned char *, unsigned char *) {}
