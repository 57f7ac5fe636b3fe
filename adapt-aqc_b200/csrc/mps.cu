// mps.cu -- MPS path of libb200aqc.so (placeholder until the MPS kernels land).
#include "ctx.h"

void b200_mps_release(b200_ctx* ctx) { (void)ctx; }
