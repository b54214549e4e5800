// Fused step driver: prepare -> gs -> finish per time step (see arb_fused.cuh).
// All three stages are lane-per-world (one world per thread); the scratch between them is tiled
// [W/32][elem][32] (arb_types.h).
#include <cuda_runtime.h>
#include <string>

#include "arb_fused.cuh"
#include "arb_internal.h"

struct FusedState {
  double* dbl = nullptr;
  int* ints = nullptr;
};

#define CUDA_OKF(call)                                                            \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      arb_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));          \
      return -100;                                                                \
    }                                                                             \
  } while (0)

#define FUSED_THREADS 64
#ifndef GS_L_SMEM
#define GS_L_SMEM 0     /* 1: the cached Lambda_FF block of the per-lane Gauss-Seidel lives in shared memory */
#endif

#ifndef PREP_MINBLOCKS
#define PREP_MINBLOCKS 1    /* resident-CTA floor of the prepare kernel: 4 = 255 registers, 6 = 168, 8 = 128 */
#endif
#ifndef GS_MINBLOCKS
#define GS_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(FUSED_THREADS, PREP_MINBLOCKS) k_fused_prepare_lane(DevModel m, DevBatch b, double dt) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= b.W) return;
  world_fused_prepare(m, fused_tile_view(b, w), w, dt);
}
__global__ void __launch_bounds__(FUSED_THREADS, GS_MINBLOCKS) k_fused_gs(DevModel m, DevBatch b, double dt) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#if GS_L_SMEM
  __shared__ double sL[36 * FUSED_THREADS];
  if (w < b.W) world_fused_gs(m, fused_tile_view(b, w), w, dt, sL + threadIdx.x, FUSED_THREADS);
#else
  double Lr[36];
  if (w < b.W) world_fused_gs(m, fused_tile_view(b, w), w, dt, Lr, 1);
#endif
}
// block-cooperative Gauss-Seidel: the sliding-friction solves of a visit are pooled over the
// block through shared memory (world_fused_gs_coop)
#define GS_COOP_THREADS 128
#define GS_COOP_CAP 48        /* queue slots: ~0.2 x 128 sliding contacts per visit in steady state, + 4 sigma */
#ifndef GS_COOP_MINBLOCKS
#define GS_COOP_MINBLOCKS 3
#endif
#define GS_COOP_SMEM (((ARB_SLIDE_NDBL + 4) * GS_COOP_CAP + 1) * sizeof(double) + \
                      (GS_COOP_CAP + 2) * sizeof(int))
__global__ void __launch_bounds__(GS_COOP_THREADS, GS_COOP_MINBLOCKS) k_fused_gs_coop(DevModel m, DevBatch b, double dt) {
  extern __shared__ double smem[];
  GsCoop co;
  co.q = smem;                                         // [21][cap]
  co.r = co.q + ARB_SLIDE_NDBL * GS_COOP_CAP;          // [4][cap]
  co.bm = (unsigned long long*)(co.r + 4 * GS_COOP_CAP);
  co.rs = (int*)(co.bm + 1);                           // [cap]
  co.cnt = co.rs + GS_COOP_CAP;                        // [2]
  co.cap = GS_COOP_CAP; co.tid = threadIdx.x; co.nthr = GS_COOP_THREADS; co.parity = 0;
  const int64_t w = (int64_t)blockIdx.x * GS_COOP_THREADS + threadIdx.x;
  const bool valid = w < b.W;
  double Lr[36];     // cached Lambda_FF block: registers (shared memory would shrink the L1)
  world_fused_gs_coop(m, fused_tile_view(b, valid ? w : b.W - 1), w, valid, dt, co, Lr, 1);
}

__global__ void __launch_bounds__(FUSED_THREADS) k_fused_finish(DevModel m, DevBatch b, double dt) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < b.W) world_fused_finish(m, fused_tile_view(b, w), w, dt);
}

// the fused stages fold controllers per dof: PD gains must be diagonal (arb_model_host.h)
bool arb_fused_supported(const arb_batch* b) { return b->model->host.fused_ok != 0; }

static int ensure_fused_scratch(arb_batch* b) {
  if (b->fused) return 0;
  FusedSizes s = fused_sizes(b->model->host);
  const int64_t W = fused_padded_worlds(b->d.W);
  FusedState* f = new FusedState();
  CUDA_OKF(cudaMalloc((void**)&f->dbl, sizeof(double) * s.total_doubles() * W));
  CUDA_OKF(cudaMalloc((void**)&f->ints, sizeof(int) * s.total_ints() * W));
  CUDA_OKF(cudaMemsetAsync(f->dbl, 0, sizeof(double) * s.total_doubles() * W, b->stream));
  CUDA_OKF(cudaMemsetAsync(f->ints, 0, sizeof(int) * s.total_ints() * W, b->stream));
  carve_fused(s, f->dbl, f->ints, b->d);
  CUDA_OKF(cudaFuncSetAttribute(k_fused_gs_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_COOP_SMEM));
  // the per-lane stages live on L1 (operands re-read every sweep / pass): no shared-memory carve-out
  CUDA_OKF(cudaFuncSetAttribute(k_fused_gs, cudaFuncAttributePreferredSharedMemoryCarveout, GS_L_SMEM ? 20 : 0));
  CUDA_OKF(cudaFuncSetAttribute(k_fused_prepare_lane, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  CUDA_OKF(cudaFuncSetAttribute(k_fused_finish, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
  b->fused = f;
  return 0;
}



int arb_fused_step(arb_batch* b, const double* dts, int nsteps) {
  int rc = ensure_fused_scratch(b);
  if (rc) return rc;
  const unsigned g = (unsigned)((b->d.W + FUSED_THREADS - 1) / FUSED_THREADS);
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (b->time_stages)
    for (int i = 0; i < 4; ++i) CUDA_OKF(cudaEventCreate(&ev[i]));
  for (int s = 0; s < nsteps; ++s) {
    const double dt = dts[s];
    if (!(dt > 0)) { arb_set_error("dt must be > 0"); return -3; }
    if (ev[0]) cudaEventRecord(ev[0], b->stream);
    k_fused_prepare_lane<<<g, FUSED_THREADS, 0, b->stream>>>(b->m, b->d, dt);
    if (ev[0]) cudaEventRecord(ev[1], b->stream);
    if (b->m.nc > 0) {
      if (b->m.nc <= 64 && b->gs_coop)
        k_fused_gs_coop<<<(unsigned)((b->d.W + GS_COOP_THREADS - 1) / GS_COOP_THREADS), GS_COOP_THREADS,
                          GS_COOP_SMEM, b->stream>>>(b->m, b->d, dt);
      else
        k_fused_gs<<<g, FUSED_THREADS, 0, b->stream>>>(b->m, b->d, dt);
    }
    if (ev[0]) cudaEventRecord(ev[2], b->stream);
    k_fused_finish<<<g, FUSED_THREADS, 0, b->stream>>>(b->m, b->d, dt);
    b->launches += (b->m.nc > 0) ? 3 : 2;
    if (ev[0]) {   // diagnostic mode: per-stage device time of this step
      cudaEventRecord(ev[3], b->stream);
      CUDA_OKF(cudaEventSynchronize(ev[3]));
      for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        b->stage_ms[i] += ms;
      }
      b->stage_ms[3] += 1.;
    }
  }
  for (int i = 0; i < 4; ++i)
    if (ev[i]) cudaEventDestroy(ev[i]);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { arb_set_error(std::string("kernel launch: ") + cudaGetErrorString(e)); return -101; }
  return 0;
}

void arb_fused_release(arb_batch* b) {
  if (!b->fused) return;
  cudaFree(b->fused->dbl);
  cudaFree(b->fused->ints);
  delete b->fused;
  b->fused = nullptr;
}
