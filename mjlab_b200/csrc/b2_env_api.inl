// b2_env_api.inl - kernels + C entry points of include/b200env.h (included by b2_api.cu).

#define B2E_WPB 4  // envs (warps) per CTA for the env kernels

__global__ void __launch_bounds__(B2E_WPB * 32) k_env_action(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KEnv e) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) env_action_world(m, d, e, w);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_env_airtime(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KEnv e, int w0, int w1) {
  int w = w0 + blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < w1) env_airtime_world(m, d, e, w);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_env_post(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KEnv e) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) env_post_world(m, d, e, w);
}
__global__ void __launch_bounds__(B2E_WPB * 32) k_env_obs(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KEnv e, int with_update) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w < d.nworld) env_obs_world(m, d, e, w, with_update);
}
// _reset_idx for the envs selected by mask (NULL = all); clears the per-env log rows first
__global__ void __launch_bounds__(B2E_WPB * 32) k_env_reset(const __grid_constant__ KModel m, const __grid_constant__ KData d, const __grid_constant__ KEnv e, const unsigned char* mask) {
  int w = blockIdx.x * B2E_WPB + (threadIdx.x >> 5);
  if (w >= d.nworld) return;
  float* logv = reinterpret_cast<float*>(e.buf[B2E_B_LOG_ENV]) + (size_t)w * B2E_NLOG;
  for (int i = threadIdx.x & 31; i < B2E_NLOG; i += 32) logv[i] = 0;
  __syncwarp();
  const bool sel = !mask || mask[w];
  if ((threadIdx.x & 31) == 0) reinterpret_cast<unsigned char*>(e.buf[B2E_B_RESET_BUF])[w] = sel;
  if (sel) env_reset_world(m, d, e, w, reinterpret_cast<const long long*>(e.buf[B2E_B_STEP_COUNTER])[0]);
  if (e.ip[B2E_IP_TERRAIN_MODE] && (threadIdx.x & 31) == 0) logv[B2E_LOG_TERRAIN_LEVEL] = (float)reinterpret_cast<const int*>(e.buf[B2E_B_TERRAIN_LEVEL])[w];
}

// One CTA per log column; fixed-order strided partial sums + shared-memory tree => deterministic.
// Column B2E_LOG_NRESET also publishes any_reset (the gate of the conditional forward); CTA 0 advances the
// env-step counter when `advance` is set (the last launch of an env step).
__global__ void __launch_bounds__(256) k_env_log(const KData d, const KEnv e, int advance) {
  __shared__ float part[256];
  const int col = blockIdx.x;
  const float* src = reinterpret_cast<const float*>(e.buf[B2E_B_LOG_ENV]);
  float acc = 0;
  for (int w = threadIdx.x; w < d.nworld; w += 256) acc += src[(size_t)w * B2E_NLOG + col];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    reinterpret_cast<float*>(e.buf[B2E_B_LOG])[col] = part[0];
    if (col == B2E_LOG_NRESET && !advance) reinterpret_cast<int*>(e.buf[B2E_B_ANY_RESET])[0] = (int)part[0];
    if (col == 0 && advance) reinterpret_cast<long long*>(e.buf[B2E_B_STEP_COUNTER])[0] += 1;
  }
}

__global__ void k_actuator_pd(float* effort, const float* pos_target, const float* vel_target, const float* effort_target,
                              const float* pos, const float* vel, const float* kp, const float* kd, const float* limit,
                              const float* sat, const float* vlim, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tau = kp[i] * (pos_target[i] - pos[i]);
  tau += kd[i] * (vel_target[i] - vel[i]);
  tau += effort_target[i];
  effort[i] = sat ? b2_dc_clip(tau, vel[i], limit[i], sat[i], vlim[i]) : fminf(fmaxf(tau, -limit[i]), limit[i]);
}

__global__ void k_delay_step(float* out, float* ring, int* count, int* head, const int* lag, const float* in, int nworld, int cap,
                             int width) {
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nworld) return;
  const int lane = threadIdx.x & 31;
  int h = head[w], c = count[w];
  // append: first push after a reset back-fills the whole ring (circular_buffer.py:190-243)
  h = (h + 1) % cap;
  float* r = ring + (size_t)w * cap * width;
  for (int i = lane; i < width; i += 32) {
    float v = in[(size_t)w * width + i];
    if (c == 0) for (int s = 0; s < cap; ++s) r[(size_t)s * width + i] = v;
    else r[(size_t)h * width + i] = v;
  }
  c = c < cap ? c + 1 : cap;
  int lg = lag[w];
  if (lg > c - 1) lg = c - 1;
  if (lg < 0) lg = 0;
  int src = (h - lg + cap) % cap;
  __syncwarp();
  for (int i = lane; i < width; i += 32) out[(size_t)w * width + i] = r[(size_t)src * width + i];
  if (lane == 0) { head[w] = h; count[w] = c; }
}

#define B2_MLP_MAXW 64
struct MlpDesc {
  const float* weight[8];
  const float* bias[8];
  int in[8], out[8];
  int nlayer, activation;
  float out_scale;
};
__device__ __forceinline__ float b2_act(float x, int kind) {
  switch (kind) {
    case 0: return x > 0 ? x : expm1f(x);          // elu
    case 1: return tanhf(x);
    case 2: return x > 0 ? x : 0.f;                 // relu
    case 3: return x / (1.f + fabsf(x));            // softsign
    default: return x;
  }
}
__global__ void k_actuator_mlp(float* effort, const float* input, const MlpDesc D, const float* limit, const float* sat,
                               const float* vlim, const float* vel, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[B2_MLP_MAXW], b[B2_MLP_MAXW];
  const int nin = D.in[0];
  for (int k = 0; k < nin; ++k) a[k] = input[i * nin + k];
  for (int l = 0; l < D.nlayer; ++l) {
    const float* W = D.weight[l];
    const float* bb = D.bias[l];
    for (int o = 0; o < D.out[l]; ++o) {
      float acc = bb ? bb[o] : 0.f;
      for (int k = 0; k < D.in[l]; ++k) acc = fmaf(W[o * D.in[l] + k], a[k], acc);
      b[o] = l + 1 < D.nlayer ? b2_act(acc, D.activation) : acc;
    }
    for (int o = 0; o < D.out[l]; ++o) a[o] = b[o];
  }
  float tau = a[0] * D.out_scale;
  if (sat) tau = b2_dc_clip(tau, vel[i], limit[i], sat[i], vlim[i]);
  else if (limit) tau = fminf(fmaxf(tau, -limit[i]), limit[i]);
  effort[i] = tau;
}

extern "C" {

int b2_sizeof_env(void) { return (int)sizeof(b2_env); }

int b2_env_check(const b2_model* m, const b2_data* d, const b2_env* e) {
  if (!m || !d || !e) return fail("null model/data/env");
  const int nj = e->ip[B2E_IP_NJ], nf = e->ip[B2E_IP_NFEET];
  if (nj < 1 || nj > 32) return fail("env: nj=%d outside [1, 32]", nj);
  if (nf < 0 || nf > B2E_MAXFEET) return fail("env: nfeet=%d exceeds %d", nf, B2E_MAXFEET);
  if (e->ip[B2E_IP_N_OBS_POLICY] != 12 + 3 * nj) return fail("env: n_obs_policy must be 12 + 3*nj");
  if (e->ip[B2E_IP_N_OBS_CRITIC] != 12 + 3 * nj + 6 * nf) return fail("env: n_obs_critic must be n_obs_policy + 6*nfeet");
  if (e->ip[B2E_IP_N_OBS_POLICY] > B2E_NDRAW - B2E_DRAW_NOISE) return fail("env: too many policy observations for the draw table");
  if (e->ip[B2E_IP_ROOT_BODY] < 0 || e->ip[B2E_IP_ROOT_BODY] >= m->nbody) return fail("env: bad root body");
  if (e->ip[B2E_IP_UPRIGHT_BODY] < 0 || e->ip[B2E_IP_UPRIGHT_BODY] >= m->nbody) return fail("env: bad upright body");
  const int adrs[3] = {e->ip[B2E_IP_LINVEL_ADR], e->ip[B2E_IP_ANGVEL_ADR], e->ip[B2E_IP_ANGMOM_ADR]};
  for (int k = 0; k < 3; ++k)
    if (adrs[k] < 0 || adrs[k] + 3 > m->nsensordata) return fail("env: sensor address %d outside sensordata", adrs[k]);
  for (int f = 0; f < nf; ++f) {
    if (e->ip[B2E_IP_FOOT_SITE + f] < 0 || e->ip[B2E_IP_FOOT_SITE + f] >= m->nsite) return fail("env: bad foot site");
    if (e->ip[B2E_IP_FOOT_FOUND_ADR + f] < 0 || e->ip[B2E_IP_FOOT_FOUND_ADR + f] >= m->nsensordata) return fail("env: bad foot found address");
    if (e->ip[B2E_IP_FOOT_FORCE_ADR + f] < 0 || e->ip[B2E_IP_FOOT_FORCE_ADR + f] + 3 > m->nsensordata) return fail("env: bad foot force address");
  }
  if (e->ip[B2E_IP_ILLEGAL_COUNT] > 0 && e->ip[B2E_IP_ILLEGAL_ADR] + e->ip[B2E_IP_ILLEGAL_COUNT] > m->nsensordata)
    return fail("env: illegal-contact sensors outside sensordata");
  for (int b = 0; b < B2E_B_COUNT; ++b)
    if (!e->buf[b] && b != B2E_B_INJECT_U) return fail("env: buffer %d is null", b);
  return 0;
}

#define B2E_LAUNCH(kern)                                                                       \
  const KModel& km = *reinterpret_cast<const KModel*>(m);                                      \
  const KData& kd = *reinterpret_cast<const KData*>(d);                                        \
  const KEnv& ke = *reinterpret_cast<const KEnv*>(e);                                          \
  kern<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke); \
  g_launches += 1;                                                                             \
  cudaError_t err = cudaGetLastError();                                                        \
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));           \
  return 0;

int b2_env_action(const b2_model* m, b2_data* d, const b2_env* e, void* stream) { B2E_LAUNCH(k_env_action) }
int b2_env_airtime(const b2_model* m, b2_data* d, const b2_env* e, void* stream) {
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  const KEnv& ke = *reinterpret_cast<const KEnv*>(e);
  k_env_airtime<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke, 0, d->nworld);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

struct AirtimeCtx { const KModel* m; const KData* d; const KEnv* e; };
static void airtime_hook(const void* ctx, cudaStream_t st, int w0, int w1) {
  const AirtimeCtx& c = *static_cast<const AirtimeCtx*>(ctx);
  k_env_airtime<<<(w1 - w0 + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, st>>>(*c.m, *c.d, *c.e, w0, w1);
  g_launches += 1;
}

int b2_env_substeps(const b2_model* m, b2_data* d, const b2_env* e, int nsub, void* stream) {
  if (nsub < 0) return fail("nsub must be >= 0");
  AirtimeCtx c{reinterpret_cast<const KModel*>(m), reinterpret_cast<const KData*>(d), reinterpret_cast<const KEnv*>(e)};
  return run_chunked(m, d, stream, nsub, 1, nullptr, nullptr, airtime_hook, &c, 1);
}

int b2_env_obs(const b2_model* m, b2_data* d, const b2_env* e, int with_update, void* stream) {
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  const KEnv& ke = *reinterpret_cast<const KEnv*>(e);
  k_env_obs<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke, with_update);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

static int launch_log(const b2_data* d, const b2_env* e, int advance, void* stream) {
  const KData& kd = *reinterpret_cast<const KData*>(d);
  const KEnv& ke = *reinterpret_cast<const KEnv*>(e);
  k_env_log<<<advance ? B2E_NLOG : 1, 256, 0, (cudaStream_t)stream>>>(kd, ke, advance);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

int b2_env_post(const b2_model* m, b2_data* d, const b2_env* e, void* stream) {
  {
    const KModel& km = *reinterpret_cast<const KModel*>(m);
    const KData& kd = *reinterpret_cast<const KData*>(d);
    const KEnv& ke = *reinterpret_cast<const KEnv*>(e);
    k_env_post<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke);
    g_launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  }
  return launch_log(d, e, 0, stream);  // column NRESET only: publishes any_reset for the conditional forward
}

int b2_env_log(const b2_data* d, const b2_env* e, void* stream) { return launch_log(d, e, 1, stream); }

int b2_env_reset(const b2_model* m, b2_data* d, const b2_env* e, const unsigned char* mask, void* stream) {
  const KModel& km = *reinterpret_cast<const KModel*>(m);
  const KData& kd = *reinterpret_cast<const KData*>(d);
  const KEnv& ke = *reinterpret_cast<const KEnv*>(e);
  k_env_reset<<<(d->nworld + B2E_WPB - 1) / B2E_WPB, B2E_WPB * 32, 0, (cudaStream_t)stream>>>(km, kd, ke, mask);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return launch_log(d, e, 0, stream);
}

int b2_actuator_pd(float* effort, const float* pos_target, const float* vel_target, const float* effort_target, const float* pos,
                   const float* vel, const float* kp, const float* kd, const float* effort_limit, const float* saturation_effort,
                   const float* velocity_limit, long long n, void* stream) {
  if (n <= 0) return 0;
  if ((saturation_effort == nullptr) != (velocity_limit == nullptr)) return fail("saturation_effort and velocity_limit go together");
  k_actuator_pd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(effort, pos_target, vel_target, effort_target, pos, vel, kp,
                                                                                  kd, effort_limit, saturation_effort, velocity_limit, n);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

int b2_delay_step(float* out, float* ring, int* count, int* head, const int* lag, const float* in, int nworld, int cap, int width,
                  void* stream) {
  if (nworld <= 0) return 0;
  if (cap < 1) return fail("delay ring capacity must be >= 1");
  k_delay_step<<<(nworld + 3) / 4, 128, 0, (cudaStream_t)stream>>>(out, ring, count, head, lag, in, nworld, cap, width);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

int b2_actuator_mlp(float* effort, const float* input, const float* const* weight, const float* const* bias, const int* layer_in,
                    const int* layer_out, int nlayer, int activation, float out_scale, const float* effort_limit,
                    const float* saturation_effort, const float* velocity_limit, const float* vel, long long n, void* stream) {
  if (n <= 0) return 0;
  if (nlayer < 1 || nlayer > 8) return fail("mlp: 1..8 layers supported, got %d", nlayer);
  MlpDesc D;
  for (int l = 0; l < nlayer; ++l) {
    if (layer_in[l] > B2_MLP_MAXW || layer_out[l] > B2_MLP_MAXW) return fail("mlp: layer width exceeds %d", B2_MLP_MAXW);
    if (l > 0 && layer_in[l] != layer_out[l - 1]) return fail("mlp: layer %d input does not match layer %d output", l, l - 1);
    D.weight[l] = weight[l]; D.bias[l] = bias ? bias[l] : nullptr; D.in[l] = layer_in[l]; D.out[l] = layer_out[l];
  }
  if (layer_out[nlayer - 1] != 1) return fail("mlp: the network must have one output");
  D.nlayer = nlayer; D.activation = activation; D.out_scale = out_scale;
  k_actuator_mlp<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(effort, input, D, effort_limit, saturation_effort,
                                                                                  velocity_limit, vel, n);
  g_launches += 1;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail("launch failed: %s", cudaGetErrorString(err));
  return 0;
}

}  // extern "C"
