// sweep_push.cu -- the fused sweep with the halo push: compute and the data movement of
// comm_isend_irecv_real (tem/source/tem_comm_module.fpp:549-646) in ONE kernel.  Elements that
// own links of the halo send buffer (prp_sendHalo, mus_construction_module.fpp:2765) store them
// into the receivers' halo rows through NVLink peer mappings right after their collision; what is
// left of the exchange is the arrival handshake (signalHaloKernel, p2p.cu).  Extra traffic in the
// sweep: one bit per element for the send mask.
#include "sweep_kernel.cuh"

namespace musb200 {

int launchSweepPush(int QQ, int relax, int kind, const SweepArgs &a, cudaStream_t st) {
  return dispatchSweep<2>(QQ, relax, kind, a, st);
}

}  // namespace musb200
