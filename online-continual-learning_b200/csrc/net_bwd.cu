// net_bwd.cu -- backward pass of the Reduced-ResNet18 engine (placeholder until the kernels land).
#include "net_ws.cuh"

extern "C" int b200ocl_net_backward(const b200ocl_net_desc*, const b200ocl_net_state*, const float*, int, void*, size_t,
                                    int, void*) {
  b200ocl::set_error("b200ocl_net_backward: not built yet");
  return B200OCL_EUNSUPPORTED;
}
