// b2_trk_api.inl - kernels + C entry points of include/b200trk.h (included by b2_api.cu).

static_assert(sizeof(KTrk) == sizeof(b2_trk), "KTrk must mirror b2_trk");

__global__ void __launch_bounds__(B2E_WPB * 32) k_trk_action(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KTrk e) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) trk_action_world(m, d, e, w);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_trk_post(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KTrk e) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) trk_post_world(m, d, e, w);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_trk_obs(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KTrk e, int with_update) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) trk_obs_world(m, d, e, w, with_update);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_trk_reset(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KTrk e, const unsigned char* mask) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w >= d.nworld) return;
  float* logv = reinterpret_cast<float*>(e.buf[B2T_B_LOG_ENV]) + (size_t)w * B2T_NLOG;
  for (int i = threadIdx.x & 31; i < B2T_NLOG; i += 32) logv[i] = 0;
  __syncwarp();
  const bool sel = !mask || mask[w];
  if ((threadIdx.x & 31) == 0) reinterpret_cast<unsigned char*>(e.buf[B2T_B_RESET_BUF])[w] = sel;
  if (sel) trk_reset_world(m, d, e, w, reinterpret_cast<const long long*>(e.buf[B2T_B_STEP_COUNTER])[0]);
}
// One CTA per log column (deterministic tree).  advance != 0 (last launch of an env step): CTA 0 also applies the adaptive
// sampler's EMA of the failed-bin counts (commands.py:405-410) and advances the env-step counter.
__global__ void __launch_bounds__(256) k_trk_log(const KData d, const __grid_constant__ KTrk e, int advance) {
  __shared__ float part[256];
  const int col = blockIdx.x;
  const float* src = reinterpret_cast<const float*>(e.buf[B2T_B_LOG_ENV]);
  float acc = 0;
  for (int w = threadIdx.x; w < d.nworld; w += 256) acc += src[(size_t)w * B2T_NLOG + col];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    reinterpret_cast<float*>(e.buf[B2T_B_LOG])[col] = part[0];
    if (col == B2T_LOG_NRESET && !advance) reinterpret_cast<int*>(e.buf[B2T_B_ANY_RESET])[0] = (int)part[0];
    if (col == 0 && advance) reinterpret_cast<long long*>(e.buf[B2T_B_STEP_COUNTER])[0] += 1;
  }
  if (col == 0 && advance && e.ip[B2T_IP_SAMPLING_MODE] == 2) {
    float* bfc = reinterpret_cast<float*>(e.buf[B2T_B_BIN_FAILED_COUNT]);
    float* cur = reinterpret_cast<float*>(e.buf[B2T_B_CURRENT_BIN_FAILED]);
    const float alpha = e.fp[B2T_FP_ADAPTIVE_ALPHA];
    for (int b = threadIdx.x; b < e.ip[B2T_IP_BIN_COUNT]; b += 256) { bfc[b] = alpha * cur[b] + (1.f - alpha) * bfc[b]; cur[b] = 0.f; }
  }
}

extern "C" {

int b2_sizeof_trk(void) { return (int)sizeof(b2_trk); }

int b2_trk_check(const b2_model* m, const b2_data* d, const b2_trk* e) {
  if (!m || !d || !e) return fail("null model/data/trk");
  const int nj = e->ip[B2T_IP_NJ], nb = e->ip[B2T_IP_NBODY];
  if (nj < 1 || nj > 32) return fail("trk: nj=%d outside [1, 32]", nj);
  if (nb < 1 || nb > B2T_MAXBODY) return fail("trk: %d tracked bodies outside [1, %d]", nb, B2T_MAXBODY);
  if (e->ip[B2T_IP_NEE] < 0 || e->ip[B2T_IP_NEE] > B2T_MAXEE) return fail("trk: too many end-effector bodies");
  if (e->ip[B2T_IP_ANCHOR_SEL] < 0 || e->ip[B2T_IP_ANCHOR_SEL] >= nb) return fail("trk: bad anchor index");
  if (e->ip[B2T_IP_N_OBS_POLICY] != 5 * nj + 15) return fail("trk: n_obs_policy must be 5*nj + 15");
  if (e->ip[B2T_IP_N_OBS_CRITIC] != 5 * nj + 15 + 9 * nb) return fail("trk: n_obs_critic must be n_obs_policy + 9*nbody");
  if (e->ip[B2T_IP_N_OBS_POLICY] > B2T_NDRAW - B2T_DRAW_NOISE) return fail("trk: too many policy observations for the draw table");
  if (e->ip[B2T_IP_TIME_STEP_TOTAL] < 2) return fail("trk: the motion needs at least 2 frames");
  if (e->ip[B2T_IP_BIN_COUNT] < 1 || e->ip[B2T_IP_KERNEL_SIZE] < 1 || e->ip[B2T_IP_KERNEL_SIZE] > B2T_MAXKERNEL) return fail("trk: bad adaptive sampler sizes");
  if (e->ip[B2T_IP_SAMPLING_MODE] < 0 || e->ip[B2T_IP_SAMPLING_MODE] > 2) return fail("trk: unknown sampling mode");
  if (e->ip[B2T_IP_ROOT_BODY] < 0 || e->ip[B2T_IP_ROOT_BODY] >= m->nbody) return fail("trk: bad root body");
  for (int b = 0; b < B2T_B_COUNT; ++b)
    if (!e->buf[b] && b != B2T_B_INJECT_U) return fail("trk: buffer %d is null", b);
  return 0;
}

#define B2T_LAUNCH(kern, ...)                                                                    \
  const KModel& km = *reinterpret_cast<const KModel*>(m);                                          \
  const KData& kd = *reinterpret_cast<const KData*>(d);                                            \
  const KTrk& ke = *reinterpret_cast<const KTrk*>(e);                                              \
  kern<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke, ##__VA_ARGS__); \
  g_launches += 1;                                                                                 \
  { cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err)); }

static int trk_launch_log(const b2_data* d, const b2_trk* e, int advance, void* stream) {
  const KData& kd = *reinterpret_cast<const KData*>(d);
  const KTrk& ke = *reinterpret_cast<const KTrk*>(e);
  k_trk_log<<<advance ? B2T_NLOG : 1, 256, 0, (cudaStream_t)stream>>>(kd, ke, advance);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

int b2_trk_action(const b2_model* m, b2_data* d, const b2_trk* e, void* stream) { B2T_LAUNCH(k_trk_action) return 0; }
int b2_trk_post(const b2_model* m, b2_data* d, const b2_trk* e, void* stream) { B2T_LAUNCH(k_trk_post) return trk_launch_log(d, e, 0, stream); }
int b2_trk_reset(const b2_model* m, b2_data* d, const b2_trk* e, const unsigned char* mask, void* stream) { B2T_LAUNCH(k_trk_reset, mask) return trk_launch_log(d, e, 0, stream); }
int b2_trk_obs(const b2_model* m, b2_data* d, const b2_trk* e, int with_update, void* stream) { B2T_LAUNCH(k_trk_obs, with_update) return 0; }
int b2_trk_log(const b2_data* d, const b2_trk* e, void* stream) { return trk_launch_log(d, e, 1, stream); }

}  // extern "C"
