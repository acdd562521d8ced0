// capi.cu -- the C ABI declared in include/btkb200.h: plan object, setup-time formulas (host, double),
// buffer management and kernel launches.  No CPU compute path exists for the hot loops: every
// analysis / beamform / synthesis / chain / covariance / solve call runs the sm_100a kernels or fails.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <cuda.h>

#include "../../include/btkb200.h"
#include "host_tables.h"
#include "launch.h"

using namespace btk;

static thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMallocHost(&p, bytes + 256);
    if (e == cudaSuccess) cap = bytes + 256;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct btkb200_plan {
  BankGeom geo;
  int C = 0, Cpad = 0, gain = 1, device = 0;
  bool has_h = false, has_g = false;
  int has_weights = 0;      // 0 none, 1 DS, 2 user/MVDR
  bool has_manifold = false;
  bool mvdr_solved = false;  // w holds solved MVDR weights: a later set_ds_weights refreshes the manifold only (the reference
                             // keeps _wmvdr apart from the quiescent vectors, beamformer.h:380-388)
  std::vector<zd> w, wq, Rn;
  std::vector<zd> ta;          // array manifold (target delay-and-sum vector): equals wq unless null-steering weights are installed
  std::vector<zd> gsc_B, gsc_wa;   // SubbandGSC: blocking matrices [B][C][C-1], active weights [B][C-1]
  bool has_gsc = false;
  std::vector<char> Rn_set;
  // device constants
  float* d_taps_h = nullptr;   // residue-major [D][TS]
  float* d_taps_g = nullptr;   // [m][M]
  cf* d_twa = nullptr;         // lane-contiguous twiddles (fb_core.cuh, FFTTables)
  cf* d_twb = nullptr;
  cf* d_wts_chain = nullptr;   // [Cpad][M]
  cf* d_w = nullptr;           // [B][C]
  cf* d_ta = nullptr;          // [B][C] array manifold (time alignment of the post-filter)
  // scratch
  DevBuf d_recs, d_work, d_in, d_out, d_aux, d_aux2, d_raw, d_adapt;
  DevBuf d_tmaps;                         // one CUtensorMap per recording of the prepared batch (warp-specialised chain)
  std::vector<long long> tmap_sig;        // batch signature + base pointer the maps were built for
  int tma_rows = 0;
  PinBuf h_desc;
  std::vector<long long> sig;  // signature of the cached chain work list
  int cached_n_work = 0;
  bool persist = false;        // warp-specialised chain: d_work holds the item prefix (WsSegs), the grid is one CTA per SM
  int item_q = 0, n_sm = 148, cta_n = 0;
  int no_prefetch = 0;         // fused chain: no L2 prefetch of the next window (chain_prepare, same footprint estimate)
  int one_cta = 0;             // fused chain: keep one CTA per SM (chain_prepare decides from the L2 footprint)
  int use_ws = 0;              // fused chain: the warp-specialised producer / consumer kernel (chain_ws.cuh) runs this shape
  int cluster = 1;             // ... with this many CTAs of a thread-block cluster sharing a work item (channel split)
  int tune_ws = -1, tune_cluster = 0;   // btkb200_plan_tune overrides (-1 / 0 = automatic)
  bool launched_chain = false;
  std::vector<int> rec_work_begin;   // first work item of every recording (+ total at the end)
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;   // copy engines of the pipelined host-buffer path
  std::vector<cudaEvent_t> ev_in, ev_k;
  long launches = 0;
  std::string err;
};

static int fail(btkb200_plan* p, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (p) p->err = buf; else g_create_error = buf;
  return code;
}

#define CK(plan, call)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(plan, BTKB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static int upload_weights(btkb200_plan* p) {
  const int M = p->geo.M, B = p->geo.B, C = p->C;
  CK(p, cudaSetDevice(p->device));
  std::vector<cf> gam;
  if (!build_chain_weight_table(p->w.data(), M, C, p->Cpad, gam)) return fail(p, BTKB200_EUNSUPPORTED, "no weight layout for M=%d", M);
  std::vector<cf> plain((size_t)B * C);
  for (size_t i = 0; i < plain.size(); i++) plain[i] = mk((float)p->w[i].real(), (float)p->w[i].imag());
  // ordered after any kernel still reading the previous tables
  CK(p, cudaStreamSynchronize(p->stream));
  CK(p, cudaMemcpy(p->d_wts_chain, gam.data(), gam.size() * sizeof(cf), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(p->d_w, plain.data(), plain.size() * sizeof(cf), cudaMemcpyHostToDevice));
  return BTKB200_OK;
}

// chunk of output frames per CTA: enough CTAs to fill 148 SMs several times over, but long enough
// that the (m R - 1)-frame warm-up of the synthesis history stays a small fraction.
static int choose_chunk(long long total_frames, int H, int W, int waves = 0) {
#ifndef BTK_CHUNK_WAVES
#define BTK_CHUNK_WAVES 2
#endif
  static const int waves_env = getenv("BTK_CHUNK_WAVES") ? atoi(getenv("BTK_CHUNK_WAVES")) : 0;   // tuning knob (A/B runs)
  const long long target_ctas = 148LL * 2 * (waves_env > 0 ? waves_env : (waves > 0 ? waves : BTK_CHUNK_WAVES));
  long long chunk = (total_frames + target_ctas - 1) / target_ctas;
  // small jobs (one short recording): prefer filling the SMs over amortising the warm-up, down to chunks whose
  // warm-up is half of their work (a 10-s recording is 139 CTAs instead of 22)
  if (chunk < H) chunk = H;
  if (chunk > 64LL * W) chunk = 64LL * W;
  // make chunk + H a multiple of W so no iteration is partly wasted
  long long its = (chunk + H + W - 1) / W;
  chunk = its * W - H;
  if (chunk < 1) chunk = W;
  return (int)chunk;
}

// Chunk size of a fused-chain batch from a list-scheduling model of the launch: CTAs are handed to the first free slot in
// launch order, a chunk of `its` iterations costs its + 0.3 (table loads, cold first window), the last chunk of a
// recording is shorter and fills gaps.  The candidate with the smallest makespan wins among those that keep at least
// two waves of CTAs in flight (a single wave measured 16 % slower than three: nothing fills the gaps behind the slowest
// CTAs; a deliberate start offset between the co-resident CTAs of the first wave made no difference, so it is not a
// lockstep effect).  The model orders the measured BTK_CHUNK_WAVES sweeps of cfg2 / cfg3 / cfg4 correctly
// (DESIGN.md); small jobs fall back to choose_chunk (fill the SMs first).
static int choose_chunk_model(const std::vector<RecDesc>& recs, int H, int W, int slots) {
  long long total = 0;
  for (size_t r = 0; r < recs.size(); r++) total += recs[r].nblk;
  int best_its = 0;
  double best = 1e300;
  std::vector<double> heap;
  for (int its = 2; its <= 64; its++) {
    const long long chunk = (long long)its * W - H;
    if (chunk < 1) continue;
    long long n_cta = 0;
    for (size_t r = 0; r < recs.size(); r++) n_cta += (recs[r].nblk + chunk - 1) / chunk;
    if (n_cta < 2LL * slots) break;            // longer chunks only make fewer CTAs
    heap.assign(slots, 0.0);                   // min-heap of slot finish times (all equal at the start)
    auto push_down = [&](size_t i) {
      for (;;) {
        size_t l = 2 * i + 1, rr = l + 1, m = i;
        if (l < heap.size() && heap[l] < heap[m]) m = l;
        if (rr < heap.size() && heap[rr] < heap[m]) m = rr;
        if (m == i) break;
        std::swap(heap[i], heap[m]); i = m;
      }
    };
    double makespan = 0.0;
    for (size_t r = 0; r < recs.size(); r++)
      for (long long j = 0; j < recs[r].nblk; j += chunk) {
        const long long nj = recs[r].nblk - j < chunk ? recs[r].nblk - j : chunk;
        const double d = (double)((nj + H + W - 1) / W) + 0.3;
        heap[0] += d;
        if (heap[0] > makespan) makespan = heap[0];
        push_down(0);
      }
    if (makespan < best) { best = makespan; best_its = its; }
  }
  if (best_its == 0) return choose_chunk(total, H, W);
  return best_its * W - H;
}

extern "C" {

int btkb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
  return n;
}

const char* btkb200_version(void) { return "btkb200 0.1 (sm_100a)"; }

const char* btkb200_last_error(const btkb200_plan* plan) { return plan ? plan->err.c_str() : g_create_error.c_str(); }

int btkb200_plan_create(btkb200_plan** out, unsigned M, unsigned m, unsigned r, unsigned dct, unsigned C,
                        const double* h, const double* g, int gain, int device) {
  if (!out) return fail(nullptr, BTKB200_EINVAL, "plan pointer is NULL");
  *out = nullptr;
  if (M == 0 || m == 0 || C == 0 || r > 8) return fail(nullptr, BTKB200_EINVAL, "bad geometry M=%u m=%u r=%u C=%u", M, m, r, C);
  const int R = 1 << r;
  // delayCompensationType 2 looks m R / 2 - 1 frames ahead (modulated.cc:285-290, an unsigned in the reference): with
  // m R = 1 that underflows and the reference's behaviour is undefined -- refused here
  if (dct == 2 && m * R < 2) return fail(nullptr, BTKB200_EINVAL, "delayCompensationType 2 needs m*R >= 2 (got %u)", m * R);
  if (!fb_supported((int)M, R) || (M % R) != 0)
    return fail(nullptr, BTKB200_EUNSUPPORTED, "no kernel for M=%u, R=%d (supported: M in {64,128,256,512,1024}, r in 0..3)", M, R);
  const int smem = fb_smem_bytes((int)M, R, (int)m);
  if (smem < 0 || smem > 227 * 1024)
    return fail(nullptr, BTKB200_EUNSUPPORTED, "prototype too long for shared memory staging: M=%u m=%u needs %d bytes", M, m, smem);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, BTKB200_ECUDA, "no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(nullptr, BTKB200_EINVAL, "device %d out of range (%d visible)", device, ndev);
  btkb200_plan* p = new btkb200_plan();
  p->geo = BankGeom((int)M, (int)m, (int)r, (int)dct);
  p->C = (int)C;
  p->Cpad = ((int)C + 3) / 4 * 4;
  p->gain = gain;
  p->device = device;
  const int N = p->geo.N, B = p->geo.B;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  std::vector<float> hf, gp(N, 0.f);
  std::vector<cf> twa, twb;
  build_analysis_taps(h, (int)M, (int)m, R, hf);
  build_fft_tables((int)M, twa, twb);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_taps_h, hf.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_taps_g, N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_twa, twa.size() * sizeof(cf));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_twb, twb.size() * sizeof(cf));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_wts_chain, (size_t)p->Cpad * M * sizeof(cf));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_w, (size_t)B * C * sizeof(cf));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_ta, (size_t)B * C * sizeof(cf));
  if (e == cudaSuccess) {
    if (h) p->has_h = true;
    if (g) { build_synthesis_taps(g, (int)M, (int)m, gp); p->has_g = true; }
    e = cudaMemcpy(p->d_taps_h, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_taps_g, gp.data(), N * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_twa, twa.data(), twa.size() * sizeof(cf), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_twb, twb.data(), twb.size() * sizeof(cf), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    fail(nullptr, BTKB200_ECUDA, "plan_create: %s", cudaGetErrorString(e));
    btkb200_plan_destroy(p);
    return BTKB200_ECUDA;
  }
  p->w.assign((size_t)B * C, zd(0, 0));
  p->wq.assign((size_t)B * C, zd(0, 0));
  p->Rn.assign((size_t)B * C * C, zd(0, 0));
  p->Rn_set.assign(B, 0);
  *out = p;
  return BTKB200_OK;
}

void btkb200_plan_destroy(btkb200_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  cudaFree(p->d_taps_h); cudaFree(p->d_taps_g); cudaFree(p->d_twa); cudaFree(p->d_twb); cudaFree(p->d_wts_chain); cudaFree(p->d_w); cudaFree(p->d_ta);
  p->d_recs.release(); p->d_work.release(); p->d_in.release(); p->d_out.release(); p->d_aux.release(); p->d_aux2.release(); p->d_raw.release(); p->d_adapt.release(); p->d_tmaps.release();
  p->h_desc.release();
  if (p->stream) cudaStreamDestroy(p->stream);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (cudaEvent_t e : p->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : p->ev_k) cudaEventDestroy(e);
  delete p;
}

int btkb200_plan_info(const btkb200_plan* p, btkb200_info* info) {
  if (!p || !info) return BTKB200_EINVAL;
  info->M = p->geo.M; info->m = p->geo.m; info->r = p->geo.r; info->R = p->geo.R; info->D = p->geo.D;
  info->N = p->geo.N; info->B = p->geo.B; info->C = p->C; info->dct = p->geo.dct;
  info->pd_analysis = p->geo.pd_a; info->pd_synthesis = p->geo.pd_s; info->laN = p->geo.laN;
  info->device = p->device; info->has_weights = p->has_weights;
  return BTKB200_OK;
}

long btkb200_nblk(const btkb200_plan* p, long T) { return p ? p->geo.nblk(T) : -1; }
long btkb200_analysis_frames(const btkb200_plan* p, long T) { return p ? p->geo.analysis_frames(T) : -1; }
long btkb200_synthesis_frames(const btkb200_plan* p, long F) { return p ? p->geo.synthesis_frames((int)F) : -1; }
long btkb200_chain_frames(const btkb200_plan* p, long T) { return p ? p->geo.chain_frames(T) : -1; }

// ------------------------------------------------------------------------------------------- weights
int btkb200_set_ds_weights(btkb200_plan* p, double fs, const double* delays, unsigned n) {
  if (!p || !delays) return BTKB200_EINVAL;
  if ((int)n != p->C)
    return fail(p, BTKB200_EINVAL, "Number of delays does not match number of channels (%u vs. %d).", n, p->C);
  ds_weights(delays, fs, p->geo.M, p->C, p->wq);
  p->ta = p->wq;
  p->has_manifold = true;
  if (!p->mvdr_solved) {
    p->w = p->wq;
    p->has_weights = 1;
  }
  {
    std::vector<cf> ta(p->wq.size());
    for (size_t i = 0; i < ta.size(); i++) ta[i] = mk((float)p->wq[i].real(), (float)p->wq[i].imag());
    CK(p, cudaSetDevice(p->device));
    CK(p, cudaStreamSynchronize(p->stream));
    CK(p, cudaMemcpy(p->d_ta, ta.data(), ta.size() * sizeof(cf), cudaMemcpyHostToDevice));
  }
  return p->mvdr_solved ? BTKB200_OK : upload_weights(p);
}

// SubbandDS::calcArrayManifoldVectors2 / calcArrayManifoldVectorsN (beamformer.cc:1100-1121 -> beamformerWeights::calcMainlobe2 /
// calcMainlobeN, :603-735): distortionless towards the target, nulls towards NC - 1 interferers.  The array manifold (what
// the post-filter time-aligns with) stays the target's delay-and-sum vector, as in the reference.
int btkb200_set_null_weights(btkb200_plan* p, double fs, const double* delaysT, unsigned n, const double* delaysJ, unsigned NC) {
  if (!p || !delaysT || !delaysJ) return BTKB200_EINVAL;
  if ((int)n != p->C)
    return fail(p, BTKB200_EINVAL, "Number of delays does not match number of channels (%u vs. %d).", n, p->C);
  if (NC < 2 || (int)NC > p->C)
    return fail(p, BTKB200_EINVAL, "1 < the number of constraints %u <= the number of sensors %d.", NC, p->C);
  std::vector<zd>& ta = p->ta;
  if (!null_weights(delaysT, delaysJ, fs, p->geo.M, p->C, (int)NC, ta, p->wq)) return fail(p, BTKB200_ESTATE, "calcNullBeamformer() failed");
  p->has_manifold = true;
  if (!p->mvdr_solved) {
    p->w = p->wq;
    p->has_weights = 1;
  }
  {
    std::vector<cf> t32(ta.size());
    for (size_t i = 0; i < ta.size(); i++) t32[i] = mk((float)ta[i].real(), (float)ta[i].imag());
    CK(p, cudaSetDevice(p->device));
    CK(p, cudaStreamSynchronize(p->stream));
    CK(p, cudaMemcpy(p->d_ta, t32.data(), t32.size() * sizeof(cf), cudaMemcpyHostToDevice));
  }
  return p->mvdr_solved ? BTKB200_OK : upload_weights(p);
}

// G1: the delay helpers of the reference's drivers, host arithmetic in the reference's own precision.
int btkb200_calc_delays_polar(float azimuth, float elevation, const double* micpos, unsigned n, double* delays) {
  if (!micpos || !delays || n == 0) return BTKB200_EINVAL;
  delays_polar2(azimuth, elevation, micpos, (int)n, delays);
  return BTKB200_OK;
}
int btkb200_calc_all_delays(double x, double y, double z, const double* micpos, unsigned n, double* delays) {
  if (!micpos || !delays || n == 0) return BTKB200_EINVAL;
  all_delays(x, y, z, micpos, (int)n, delays);
  return BTKB200_OK;
}

int btkb200_get_array_manifold(const btkb200_plan* p, double* w) {
  if (!p || !w) return BTKB200_EINVAL;
  if (!p->has_manifold || p->ta.empty()) return fail(const_cast<btkb200_plan*>(p), BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  for (size_t i = 0; i < p->ta.size(); i++) { w[2 * i] = p->ta[i].real(); w[2 * i + 1] = p->ta[i].imag(); }
  return BTKB200_OK;
}

int btkb200_set_weights(btkb200_plan* p, const double* w) {
  if (!p || !w) return BTKB200_EINVAL;
  for (size_t i = 0; i < p->w.size(); i++) p->w[i] = zd(w[2 * i], w[2 * i + 1]);
  p->has_weights = 2;
  p->mvdr_solved = false;
  return upload_weights(p);
}

int btkb200_get_weights(const btkb200_plan* p, double* w) {
  if (!p || !w) return BTKB200_EINVAL;
  if (!p->has_weights) return fail(const_cast<btkb200_plan*>(p), BTKB200_ESTATE, "no weights installed");
  for (size_t i = 0; i < p->w.size(); i++) { w[2 * i] = p->w[i].real(); w[2 * i + 1] = p->w[i].imag(); }
  return BTKB200_OK;
}

int btkb200_get_manifold(const btkb200_plan* p, double* w) {
  if (!p || !w) return BTKB200_EINVAL;
  if (!p->has_manifold) return fail(const_cast<btkb200_plan*>(p), BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  for (size_t i = 0; i < p->wq.size(); i++) { w[2 * i] = p->wq[i].real(); w[2 * i + 1] = p->wq[i].imag(); }
  return BTKB200_OK;
}

// ------------------------------------------------------------------------------------------- prototype design
static int design_common(int kind, const double* h, unsigned M, unsigned m, unsigned r, double v, double wp_factor, int tau,
                         double tolerance, int device, double* out, double* err) {
  if (!out || M == 0 || m == 0 || r > 8 || (M >> r) == 0 || ((M >> r) << r) != M || !(wp_factor > 0.0) || !(tolerance >= 0.0))
    return fail(nullptr, BTKB200_EINVAL, "bad design arguments M=%u m=%u r=%u", M, m, r);
  if ((M * m) & 1u) return fail(nullptr, BTKB200_EUNSUPPORTED, "prototype length M*m = %u must be even", M * m);
  if ((size_t)M * m > 8192) return fail(nullptr, BTKB200_EUNSUPPORTED, "prototype length M*m = %u too large (max 8192)", M * m);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, BTKB200_ECUDA, "no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(nullptr, BTKB200_EINVAL, "device %d out of range (%d visible)", device, ndev);
  cudaError_t e = cudaSetDevice(device);
  int sweeps = 0;
  if (e == cudaSuccess) e = design_prototype(kind, h, (int)M, (int)m, (int)r, v, wp_factor, tau, tolerance, out, err, &sweeps);
  if (e != cudaSuccess) return fail(nullptr, BTKB200_ECUDA, "prototype design failed: %s", cudaGetErrorString(e));
  return BTKB200_OK;
}

int btkb200_design_analysis_prototype(unsigned M, unsigned m, unsigned r, double wp_factor, int tau, double tolerance, int device,
                                      double* h, double* err) {
  return design_common(0, nullptr, M, m, r, 0.0, wp_factor, tau, tolerance, device, h, err);
}

int btkb200_design_synthesis_prototype(const double* h, unsigned M, unsigned m, unsigned r, double v, double wp_factor, int tau,
                                       double tolerance, int device, double* g, double* err) {
  if (!h) return fail(nullptr, BTKB200_EINVAL, "analysis prototype is NULL");
  return design_common(1, h, M, m, r, v, wp_factor, tau, tolerance, device, g, err);
}

static int design_nyquist_common(int kind, const double* h, unsigned M, unsigned m, unsigned r, double wp_factor, int tau,
                                 double tolerance, int device, double* out, int* path) {
  if (!out || M == 0 || m == 0 || r > 8 || (M >> r) == 0 || ((M >> r) << r) != M || !(wp_factor > 0.0) || !(tolerance > 0.0))
    return fail(nullptr, BTKB200_EINVAL, "bad design arguments M=%u m=%u r=%u", M, m, r);
  if ((size_t)M * m > 4096) return fail(nullptr, BTKB200_EUNSUPPORTED, "prototype length M*m = %u too large (max 4096)", M * m);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, BTKB200_ECUDA, "no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(nullptr, BTKB200_EINVAL, "device %d out of range (%d visible)", device, ndev);
  cudaError_t e = cudaSetDevice(device);
  int sweeps = 0;
  if (e == cudaSuccess) e = design_prototype_nyquist(kind, h, (int)M, (int)m, (int)r, wp_factor, tau, tolerance, out, path, &sweeps);
  if (e != cudaSuccess) return fail(nullptr, BTKB200_ECUDA, "prototype design failed: %s", cudaGetErrorString(e));
  return BTKB200_OK;
}

int btkb200_design_analysis_nyquist(unsigned M, unsigned m, unsigned r, double wp_factor, int tau, double tolerance, int device,
                                    double* h, int* path) {
  return design_nyquist_common(0, nullptr, M, m, r, wp_factor, tau, tolerance, device, h, path);
}

int btkb200_design_synthesis_nyquist(const double* h, unsigned M, unsigned m, unsigned r, double wp_factor, int tau, double tolerance,
                                     int device, double* g, int* path) {
  if (!h) return fail(nullptr, BTKB200_EINVAL, "analysis prototype is NULL");
  return design_nyquist_common(1, h, M, m, r, wp_factor, tau, tolerance, device, g, path);
}

// ------------------------------------------------------------------------------------------- GSC (fixed active weights)
int btkb200_gsc_calc_weights(btkb200_plan* p, double fs, const double* delays, unsigned n) {
  if (!p || !delays) return BTKB200_EINVAL;
  if (p->C <= 1) return fail(p, BTKB200_EINVAL, "The number of channels must be > 1 but it is %d", p->C);   // beamformer.cc:536-539
  int rc = btkb200_set_ds_weights(p, fs, delays, n);
  if (rc) return rc;
  const int B = p->geo.B, C = p->C, bs = C - 1;
  p->gsc_B.assign((size_t)B * C * bs, zd(0, 0));
  p->gsc_wa.assign((size_t)B * bs, zd(0, 0));
  std::vector<zd> Bm;
  for (int s = 0; s < B; s++) {
    if (!blocking_matrix(&p->wq[(size_t)s * C], C, 1, Bm)) return fail(p, BTKB200_EINVAL, "_calcBlockingMatrix() failed");
    std::copy(Bm.begin(), Bm.end(), p->gsc_B.begin() + (size_t)s * C * bs);
  }
  p->has_gsc = true;
  return BTKB200_OK;
}

int btkb200_gsc_set_active_weights(btkb200_plan* p, unsigned bin, const double* packed, unsigned n) {
  if (!p || !packed) return BTKB200_EINVAL;
  if (!p->has_gsc) return fail(p, BTKB200_ESTATE, "call calcGSCWeightsX() once");          // beamformer.cc:1427-1430
  const unsigned bs = (unsigned)p->C - 1;
  if (n != 2 * bs)
    return fail(p, BTKB200_EINVAL, "the size of an active weight vector must be %u but it is %u", 2 * bs, n);   // :764-766
  if ((int)bin >= p->geo.B) return fail(p, BTKB200_EINVAL, "Must be a frequency bin %u < %d", bin, p->geo.B);
  for (unsigned k = 0; k < bs; k++) p->gsc_wa[(size_t)bin * bs + k] = zd(packed[2 * k], packed[2 * k + 1]);
  return BTKB200_OK;
}

int btkb200_gsc_zero_active_weights(btkb200_plan* p) {
  if (!p) return BTKB200_EINVAL;
  if (!p->has_gsc) return fail(p, BTKB200_ESTATE, "call calcGSCWeightsX() once");
  std::fill(p->gsc_wa.begin(), p->gsc_wa.end(), zd(0, 0));
  return BTKB200_OK;
}

int btkb200_gsc_get_blocking_matrix(const btkb200_plan* p, unsigned bin, double* Bout) {
  if (!p || !Bout || (int)bin >= p->geo.B) return BTKB200_EINVAL;
  if (!p->has_gsc) return fail(const_cast<btkb200_plan*>(p), BTKB200_ESTATE, "call calcGSCWeightsX() once");
  const size_t n = (size_t)p->C * (p->C - 1);
  for (size_t i = 0; i < n; i++) { Bout[2 * i] = p->gsc_B[bin * n + i].real(); Bout[2 * i + 1] = p->gsc_B[bin * n + i].imag(); }
  return BTKB200_OK;
}

int btkb200_gsc_apply(btkb200_plan* p, int normalize) {
  if (!p) return BTKB200_EINVAL;
  if (!p->has_gsc) return fail(p, BTKB200_ESTATE, "call calcGSCWeightsX() once");
  const int B = p->geo.B, C = p->C, bs = C - 1;
  p->w = p->wq;                       // bin 0: the quiescent vector alone (beamformer.cc:1320-1324)
  for (int s = 1; s < B; s++) {
    double nn = 0;
    for (int c = 0; c < C; c++) {
      zd wl(0, 0);                    // (B wa)_c   (zgemv, beamformer.cc:782)
      for (int k = 0; k < bs; k++) wl += p->gsc_B[((size_t)s * C + c) * bs + k] * p->gsc_wa[(size_t)s * bs + k];
      const zd v = p->wq[(size_t)s * C + c] - wl;
      p->w[(size_t)s * C + c] = v;
      nn += std::norm(v);
    }
    if (normalize) {                  // w <- w / (||w|| C)   (calcOutputOfGSC, beamformer.cc:1273-1282)
      const double d = sqrt(nn) * C;
      for (int c = 0; c < C; c++) p->w[(size_t)s * C + c] /= d;
    }
  }
  p->has_weights = 2;
  return upload_weights(p);
}

// ------------------------------------------------------------------------------------------- MVDR setup
int btkb200_set_covariance(btkb200_plan* p, unsigned bin, const double* R, unsigned rows, unsigned cols) {
  if (!p || !R) return BTKB200_EINVAL;
  if ((int)rows != p->C || (int)cols != p->C)
    return fail(p, BTKB200_EINVAL, "The spatial spectral matrix must be %d x %d but it is %u x %u", p->C, p->C, rows, cols);
  if ((int)bin >= p->geo.B) return fail(p, BTKB200_EINVAL, "bin %u out of range (0..%d)", bin, p->geo.B - 1);
  const size_t CC = (size_t)p->C * p->C;
  for (size_t i = 0; i < CC; i++) p->Rn[bin * CC + i] = zd(R[2 * i], R[2 * i + 1]);
  p->Rn_set[bin] = 1;
  return BTKB200_OK;
}

int btkb200_get_covariance(const btkb200_plan* p, unsigned bin, double* R) {
  if (!p || !R || (int)bin >= p->geo.B) return BTKB200_EINVAL;
  if (!p->Rn_set[bin]) return fail(const_cast<btkb200_plan*>(p), BTKB200_ESTATE, "no spatial spectral matrix for bin %u", bin);
  const size_t CC = (size_t)p->C * p->C;
  for (size_t i = 0; i < CC; i++) { R[2 * i] = p->Rn[bin * CC + i].real(); R[2 * i + 1] = p->Rn[bin * CC + i].imag(); }
  return BTKB200_OK;
}

int btkb200_set_diffuse_noise_model(btkb200_plan* p, const double* micpos, unsigned n_mics, double fs, double sspeed) {
  if (!p || !micpos) return BTKB200_EINVAL;
  if ((int)n_mics != p->C)
    return fail(p, BTKB200_EINVAL, "The number of microphones must be %d but it is %u", p->C, n_mics);
  diffuse_model(micpos, p->C, fs, sspeed, p->geo.M, p->Rn);
  std::fill(p->Rn_set.begin(), p->Rn_set.end(), 1);
  return BTKB200_OK;
}

int btkb200_diag_load_bin(btkb200_plan* p, unsigned bin, float w) {
  if (!p || (int)bin >= p->geo.B) return BTKB200_EINVAL;
  if (!p->Rn_set[bin]) return fail(p, BTKB200_ESTATE, "Construct first a noise covariance matrix");
  const size_t C = p->C;
  // the weight is stored as float before it is added (beamformer.cc:2342, 2562-2565)
  for (size_t c = 0; c < C; c++) p->Rn[(bin * C + c) * C + c] += zd((double)w, 0.0);
  return BTKB200_OK;
}

int btkb200_diag_load(btkb200_plan* p, float w) {
  if (!p) return BTKB200_EINVAL;
  for (int s = 0; s < p->geo.B; s++) {
    int rc = btkb200_diag_load_bin(p, (unsigned)s, w);
    if (rc != BTKB200_OK) return rc;
  }
  return BTKB200_OK;
}

int btkb200_divide_nondiagonal(btkb200_plan* p, float mu) {
  if (!p) return BTKB200_EINVAL;
  const size_t C = p->C;
  const double den = 1.0 + (double)mu;   // (1.0+myu) with a float myu, beamformer.h:371-376
  for (int s = 0; s < p->geo.B; s++) {
    if (!p->Rn_set[s]) return fail(p, BTKB200_ESTATE, "Construct first a noise covariance matrix");
    for (size_t a = 0; a < C; a++)
      for (size_t b = 0; b < C; b++)
        if (a != b) p->Rn[((size_t)s * C + a) * C + b] /= den;
  }
  return BTKB200_OK;
}

int btkb200_solve_mvdr(btkb200_plan* p, double /*sample_rate: unused by the reference too*/, double dThreshold,
                       int* n_fallback) {
  if (!p) return BTKB200_EINVAL;
  if (!p->Rn_set[0]) return fail(p, BTKB200_ESTATE, "Set a spatial spectral matrix before calling calcMVDRWeights()");
  if (!p->has_manifold) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  const int B = p->geo.B, C = p->C;
  for (int s = 1; s < B; s++)
    if (!p->Rn_set[s]) return fail(p, BTKB200_ESTATE, "no spatial spectral matrix for bin %d", s);
  if (C > 64) return fail(p, BTKB200_EUNSUPPORTED, "MVDR solve supports at most 64 channels (got %d)", C);
  CK(p, cudaSetDevice(p->device));
  const size_t nR = (size_t)B * C * C, nW = (size_t)B * C;
  CK(p, p->d_aux.reserve(nR * sizeof(double2) + 2 * nW * sizeof(double2) + B * sizeof(int)));
  double2* dR = (double2*)p->d_aux.p;
  double2* dd = dR + nR;
  double2* dw = dd + nW;
  int* dfb = (int*)(dw + nW);
  CK(p, cudaMemcpyAsync(dR, p->Rn.data(), nR * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
  CK(p, cudaMemcpyAsync(dd, p->wq.data(), nW * sizeof(double2), cudaMemcpyHostToDevice, p->stream));
  CK(p, launch_mvdr_solve(dR, dd, dw, dfb, B, C, dThreshold, p->stream));
  p->launches++;
  std::vector<zd> wnew(nW);
  std::vector<int> fb(B);
  CK(p, cudaMemcpyAsync(wnew.data(), dw, nW * sizeof(double2), cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaMemcpyAsync(fb.data(), dfb, B * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  int nfb = 0;
  for (int s = 0; s < B; s++) nfb += fb[s];
  if (n_fallback) *n_fallback = nfb;
  p->w = wnew;
  p->has_weights = 2;
  p->mvdr_solved = true;
  return upload_weights(p);
}

// ------------------------------------------------------------------------------------------- staged (device)
static int upload_desc(btkb200_plan* p, const std::vector<RecDesc>& recs, const std::vector<WorkItem>& work,
                       cudaStream_t st) {
  const size_t br = recs.size() * sizeof(RecDesc), bw = work.size() * sizeof(WorkItem);
  CK(p, p->d_recs.reserve(br));
  CK(p, p->d_work.reserve(bw));
  CK(p, p->h_desc.reserve(br + bw));
  memcpy(p->h_desc.p, recs.data(), br);
  memcpy((char*)p->h_desc.p + br, work.data(), bw);
  CK(p, cudaMemcpyAsync(p->d_recs.p, p->h_desc.p, br, cudaMemcpyHostToDevice, st));
  CK(p, cudaMemcpyAsync(p->d_work.p, (char*)p->h_desc.p + br, bw, cudaMemcpyHostToDevice, st));
  // the pinned staging area is reused by the next call: make sure the copies have left it
  CK(p, cudaStreamSynchronize(st));
  return BTKB200_OK;
}

int btkb200_analysis_dev(btkb200_plan* p, const float* d_pcm, long T, float* d_snap, void* stream) {
  if (!p || !d_pcm || !d_snap || T < 0) return BTKB200_EINVAL;
  if (!p->has_h) return fail(p, BTKB200_ESTATE, "plan was created without an analysis prototype");
  CK(p, cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int F = p->geo.analysis_frames(T);
  std::vector<RecDesc> recs(1);
  recs[0].pcm_off = 0; recs[0].out_off = 0; recs[0].T = (int)T; recs[0].nblk = F;
  std::vector<WorkItem> work;
  const int W = fb_frames_per_iter(p->geo.M, p->geo.R);
  long long chunk = ((long long)F + 148 * 4 - 1) / (148 * 4);
  chunk = (chunk + W - 1) / W * W;
  if (chunk < W) chunk = W;
  build_work(recs, (int)chunk, work);
  p->sig.clear();
  int rc = upload_desc(p, recs, work, st);
  if (rc) return rc;
  AnalysisParams a;
  a.pcm = d_pcm; a.snap = (cf*)d_snap; a.recs = (const RecDesc*)p->d_recs.p; a.work = (const WorkItem*)p->d_work.p;
  a.taps_h = p->d_taps_h; a.twa = p->d_twa; a.twb = p->d_twb; a.C = p->C; a.Cpad = p->Cpad; a.m = p->geo.m; a.laN = p->geo.laN;
  // channels are independent in the analysis bank: when the frame chunks alone do not fill the SMs, the channel
  // groups of a chunk are spread over several CTAs
  const int n_cg = p->Cpad / 4;
  int slices = work.empty() ? 1 : (int)(2 * 148 / work.size());
  if (slices < 1) slices = 1;
  if (slices > n_cg) slices = n_cg;
  a.cg_slices = slices;
  if (!work.empty()) { CK(p, launch_analysis(p->geo.M, p->geo.R, a, (int)work.size() * slices, st)); p->launches++; }
  return BTKB200_OK;
}

int btkb200_beamform_dev(btkb200_plan* p, const float* d_snap, long F, float* d_Y, void* stream) {
  if (!p || !d_snap || !d_Y || F < 0) return BTKB200_EINVAL;
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  CK(p, cudaSetDevice(p->device));
  if (F > 0) { CK(p, launch_beamform((const cf*)d_snap, p->d_w, (cf*)d_Y, F, p->geo.B, p->C, (cudaStream_t)stream)); p->launches++; }
  return BTKB200_OK;
}

int btkb200_beamform_zelinski_dev(btkb200_plan* p, const float* d_snap, long F, double alpha, int type, int min_frames,
                                  float* d_Y, float* d_W, void* stream) {
  if (!p || !d_snap || !d_Y || F < 0) return BTKB200_EINVAL;
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  if (!p->has_manifold) return fail(p, BTKB200_ESTATE, "set beamformer's weights");
  if (p->C <= 1) return fail(p, BTKB200_EINVAL, "The number of channels %d is <= 1", p->C);   // postfilter.cc:62-65
  CK(p, cudaSetDevice(p->device));
  if (F == 0) return BTKB200_OK;
  CK(p, p->d_aux.reserve(zelinski_scratch_bytes(F, p->geo.B)));
  CK(p, launch_beamform_zelinski((const cf*)d_snap, p->d_w, p->d_ta, (cf*)d_Y, (float4*)p->d_aux.p, d_W, F, p->geo.B, p->C,
                                 alpha, type, min_frames, (cudaStream_t)stream));
  p->launches += 4;
  return BTKB200_OK;
}

int btkb200_synthesis_dev(btkb200_plan* p, const float* d_Y, long F, float* d_out, void* stream) {
  if (!p || !d_Y || !d_out || F < 0) return BTKB200_EINVAL;
  if (!p->has_g) return fail(p, BTKB200_ESTATE, "plan was created without a synthesis prototype");
  CK(p, cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nout = p->geo.synthesis_frames((int)F);
  if (nout == 0) return BTKB200_OK;
  std::vector<RecDesc> recs(1);
  recs[0].pcm_off = 0; recs[0].out_off = 0; recs[0].T = (int)F; recs[0].nblk = nout;
  std::vector<WorkItem> work;
  const int W = fb_frames_per_iter(p->geo.M, p->geo.R), H = p->geo.m * p->geo.R - 1;
  build_work(recs, choose_chunk(nout, H, W), work);
  p->sig.clear();
  int rc = upload_desc(p, recs, work, st);
  if (rc) return rc;
  SynthesisParams s;
  s.Y = (const cf*)d_Y; s.out = d_out; s.recs = (const RecDesc*)p->d_recs.p; s.work = (const WorkItem*)p->d_work.p;
  s.taps_g = p->d_taps_g; s.twa = p->d_twa; s.twb = p->d_twb; s.m = p->geo.m; s.pd_s = p->geo.pd_s; s.gain = p->gain;
  CK(p, launch_synthesis(p->geo.M, p->geo.R, s, (int)work.size(), st));
  p->launches++;
  return BTKB200_OK;
}

// ------------------------------------------------------------------------------------------- fused (device)
// Build (or reuse) the work list of a batch: every recording is cut into chunks of output frames, one CTA each.
static int chain_prepare(btkb200_plan* p, const long long* pcm_off, const long long* T, const long long* out_off, int n,
                         cudaStream_t st) {
  std::vector<long long> sig;
  sig.reserve(3 * (size_t)n + 1);
  sig.push_back(n);
  for (int i = 0; i < n; i++) { sig.push_back(pcm_off[i]); sig.push_back(T[i]); sig.push_back(out_off[i]); }
  if (sig == p->sig) return BTKB200_OK;
  std::vector<RecDesc> recs(n);
  long long total = 0;
  for (int i = 0; i < n; i++) {
    if (T[i] < 0 || T[i] > 0x7ffffff0LL) return fail(p, BTKB200_EINVAL, "recording %d: bad length %lld", i, T[i]);
    recs[i].pcm_off = pcm_off[i]; recs[i].out_off = out_off[i]; recs[i].T = (int)T[i]; recs[i].nblk = p->geo.chain_frames(T[i]);
    total += recs[i].nblk;
  }
  std::vector<WorkItem> work;
  const int H = p->geo.m * p->geo.R - 1;
  int l2 = 0;
  CK(p, cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, p->device));
  // The warp-specialised kernel (one CTA per SM: eight transform warps fed by a producer warpgroup through two
  // mbarrier-guarded stages) runs every shape whose stages fit shared memory; BTK_CHAIN_WS=0 keeps the first
  // sessions' kernel for A/B runs.
  {
    static const int ws_env = getenv("BTK_CHAIN_WS") ? atoi(getenv("BTK_CHAIN_WS")) : -1;
    const int Wws = chain_ws_frames_per_iter(p->geo.M, p->geo.R, p->geo.m);
    // automatic choice: from M = 128 up (M = 128: 0.2711 ms against 0.2727 ms for the first sessions' kernel with two CTAs per
    // SM once the tensor-copy producer feeds it; M = 64 stays with the old kernel)
    p->use_ws = (Wws > 0 && (p->tune_ws >= 0 ? p->tune_ws != 0 : (ws_env >= 0 ? ws_env != 0 : p->geo.M >= 128))) ? 1 : 0;
    p->cluster = 1;
    if (p->use_ws) {
      // Channel split over a thread-block cluster: S CTAs share a work item, each stages only its own channel groups and
      // the partial beamformer outputs are summed through distributed shared memory.  Without it one CTA walks the rows
      // of interleaved PCM (all channels per row) Cpad/4 times out of L2, and the windows of 148 CTAs have to survive
      // there: 111 MB at 64 channels (M = 512), 189 MB for M = 256.  The split costs a rendezvous and two
      // exchanges per iteration while the transforms per CTA shrink by S.  With the tensor-copy producer (which fetches
      // full 128-byte lines into L2 and keeps the load/store unit out of it) the unsplit kernel stays ahead until the
      // windows in flight exceed L2 itself: measured M = 512, 64 channels (111 MB) 1.25 ms at S = 1, 1.36 at 2, 1.91 at 4;
      // M = 256, 64 channels (189 MB) 0.473 / 0.427 / 0.505 ms; 32 channels (95 MB) 0.403 / 0.454 ms.  So: the smallest power
      // of two that brings the windows in flight under 1.2 x L2.
      static const int cl_env = getenv("BTK_CLUSTER") ? atoi(getenv("BTK_CLUSTER")) : 0;
      const int n_groups = p->Cpad / 4;
      const double window_all = (double)(Wws - 1 + p->geo.m * p->geo.R) * p->geo.D * p->Cpad * sizeof(float);
      int S = 1;
      while (S < 8 && n_groups % (2 * S) == 0 && chain_ws_cluster_ok(p->geo.M, p->geo.R, p->geo.m, 2 * S) &&
             148.0 * window_all / S > 1.2 * l2)
        S *= 2;
      if (cl_env > 0 && n_groups % cl_env == 0 && chain_ws_cluster_ok(p->geo.M, p->geo.R, p->geo.m, cl_env)) S = cl_env;
      if (p->tune_cluster > 0) S = p->tune_cluster;     // validated by btkb200_plan_tune
      p->cluster = S;
    }
    p->persist = false;
    if (p->use_ws) {
      // Persistent schedule (chain_ws.cuh::WsSegs): one CTA (cluster) per SM, each takes a contiguous share of the batch's
      // iterations -- no wave quantisation (cfg3: 640 chunk CTAs were 4.32 waves), one table load and one cold stage per
      // SM, one history warm-up per (CTA, recording) instead of one per chunk.  BTK_WS_PERSIST=0 keeps the chunk list.
      static const int persist_env = getenv("BTK_WS_PERSIST") ? atoi(getenv("BTK_WS_PERSIST")) : 1;
      p->one_cta = 1; p->no_prefetch = 1;
      // (with a channel-split cluster the chunk list stays: 64 channels at M = 256, cluster of 2, measured 0.423 ms with
      // chunks in several waves against 0.479 ms with one persistent cluster per SM pair)
      if (persist_env && p->cluster == 1 && !getenv("BTK_CHUNK_WAVES")) {
        CK(p, cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, p->device));
        // items of W / 8 frames; for the launch of the whole batch the CTAs' first items are balanced by ITERATIONS
        // (host_tables.h::balance_ctas), launches of a recording range split their items evenly
        const int q = Wws >= 8 ? Wws / 8 : 1;
        std::vector<int> prefix(n + 1, 0);
        for (int i = 0; i < n; i++) prefix[i + 1] = prefix[i] + (recs[i].nblk + q - 1) / q;
        static const int bal_env = getenv("BTK_WS_BALANCE") ? atoi(getenv("BTK_WS_BALANCE")) : 1;   // A/B runs
        p->cta_n = 0;
        if (bal_env && prefix[n] > 0) {
          // only where it saves an iteration: with hundreds of iterations per CTA the even split is as short, and it
          // measured 4 % FASTER there (1 024 utterances of 64 channels: 31.6 against 32.9 ms)
          std::vector<int> begin;
          const long long B = balance_ctas(recs, prefix, q, Wws, H, p->n_sm, begin);
          if (B < even_split_iterations(recs, prefix, q, Wws, H, p->n_sm)) {
            p->cta_n = p->n_sm;
            prefix.insert(prefix.end(), begin.begin(), begin.end());   // uploaded behind the prefix sums
          }
        }
        const size_t br = recs.size() * sizeof(RecDesc), bw = prefix.size() * sizeof(int);
        CK(p, p->d_recs.reserve(br));
        CK(p, p->d_work.reserve(bw));
        CK(p, p->h_desc.reserve(br + bw));
        memcpy(p->h_desc.p, recs.data(), br);
        memcpy((char*)p->h_desc.p + br, prefix.data(), bw);
        CK(p, cudaMemcpyAsync(p->d_recs.p, p->h_desc.p, br, cudaMemcpyHostToDevice, st));
        CK(p, cudaMemcpyAsync(p->d_work.p, (char*)p->h_desc.p + br, bw, cudaMemcpyHostToDevice, st));
        CK(p, cudaStreamSynchronize(st));
        p->rec_work_begin.assign(prefix.begin(), prefix.begin() + n + 1);   // callers launch recording ranges [r0, r1) as item ranges
        p->persist = true;
        p->item_q = q;
        p->sig = sig;
        p->cached_n_work = p->rec_work_begin[n];
        return BTKB200_OK;
      }
      int slots = 148 / p->cluster;
      if (getenv("BTK_CHUNK_WAVES")) build_work(recs, choose_chunk(total, H, Wws), work);
      else build_work(recs, choose_chunk_model(recs, H, Wws, slots), work);
    }
  }
  const int W = chain_frames_per_iter(p->geo.M, p->geo.R, p->geo.m);
  if (!p->use_ws) {
  // two CTAs per SM for M <= 256 (kern_fb.cuh KernCfg::MINB), one otherwise.  Every resident CTA walks its window of
  // (W - 1 + mR) blocks of D time steps once per group of four channels, and the rows of interleaved PCM it touches hold
  // ALL channels: when the windows of all resident CTAs together no longer fit L2, each of the Cpad/4 passes goes back to
  // HBM (measured: M=256, 32 channels 2.43 ms with two CTAs per SM, 1.50 ms with one; 64 channels 5.23 -> 2.94 ms; below
  // ~100 MB two CTAs stay ahead).  Then keep one CTA per SM.
  int cps = p->geo.M <= 256 ? 2 : 1;
  const double window_bytes = (double)(W - 1 + p->geo.m * p->geo.R) * p->geo.D * p->Cpad * sizeof(float);
  if (cps == 2) {
    static const int one_cta_env = getenv("BTK_ONE_CTA") ? atoi(getenv("BTK_ONE_CTA")) : -1;   // A/B runs
    p->one_cta = one_cta_env >= 0 ? one_cta_env : (148.0 * cps * window_bytes > 0.8 * l2);
    if (p->one_cta) cps = 1;
  }
  {
    static const int npf_env = getenv("BTK_NO_PREFETCH") ? atoi(getenv("BTK_NO_PREFETCH")) : -1;   // A/B runs
    // the L2 prefetch of the next window's new rows (all channels) evicts rows still in use once a row of PCM is a full
    // 128-byte line or more: without it -1..-2 % at 32 channels (M = 128 .. 512), -2 % (M=512) to -5 % (M=1024) at 64,
    // but +0.4..2.4 % at 16 channels, whatever the footprint -- the channel count decides, not the bytes in flight
    p->no_prefetch = npf_env >= 0 ? npf_env : (p->Cpad >= 32);
  }
  // BTK_CHUNK_WAVES overrides the chunk model (A/B runs)
  if (getenv("BTK_CHUNK_WAVES")) build_work(recs, choose_chunk(total, H, W), work);
  else build_work(recs, choose_chunk_model(recs, H, W, 148 * cps), work);
  }
  p->rec_work_begin.assign(n + 1, 0);
  for (size_t w = 0; w < work.size(); w++) p->rec_work_begin[work[w].rec + 1]++;
  for (int i = 0; i < n; i++) p->rec_work_begin[i + 1] += p->rec_work_begin[i];
  int rc = upload_desc(p, recs, work, st);
  if (rc) return rc;
  p->sig = sig;
  p->cached_n_work = (int)work.size();
  return BTKB200_OK;
}

// ---- tensor maps of the prepared batch (warp-specialised chain with the raw window layout) -----------------------------
// Recording i of the batch is described to the tensor copy engine as a 2-D float tensor [T_i][C] at d_pcm + pcm_off_i with
// boxes of (4 channels x rows time steps).  One map per recording: time steps outside [0, T_i) are then outside the tensor
// and come back as zeros -- the zero history and zero tail of the analysis bank -- instead of the neighbouring recording.
typedef CUresult (*btk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static btk_encode_tiled_fn tensor_map_encoder() {
  static btk_encode_tiled_fn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return (btk_encode_tiled_fn)ptr;
  }();
  return fn;
}

// Returns the device array of maps, or NULL when this batch goes through the register-load producer (channel count not a
// multiple of 4, unaligned buffer, BTK_WS_TMA=0).
static const void* chain_tensor_maps(btkb200_plan* p, const float* d_pcm, cudaStream_t st) {
  static const int tma_env = getenv("BTK_WS_TMA") ? atoi(getenv("BTK_WS_TMA")) : 1;
  if (!BTK_WS_RAW || !tma_env || !p->use_ws || p->C % 4 != 0 || (reinterpret_cast<uintptr_t>(d_pcm) & 15) != 0 || p->sig.empty()) return nullptr;
  btk_encode_tiled_fn enc = tensor_map_encoder();
  if (!enc) return nullptr;
  const int n = (int)p->sig[0];
  for (int i = 0; i < n; i++) if (p->sig[1 + 3 * i] % 4 != 0) return nullptr;
  std::vector<long long> key(p->sig);
  key.push_back((long long)reinterpret_cast<uintptr_t>(d_pcm));
  if (key == p->tmap_sig && p->d_tmaps.p) return p->d_tmaps.p;
  int rows = p->geo.D < 256 ? p->geo.D : 256;
  // A/B knobs (read once): box height and L2 promotion of the maps
  static const int rows_env = getenv("BTK_TMA_ROWS") ? atoi(getenv("BTK_TMA_ROWS")) : 0;
  static const int promo_env = getenv("BTK_TMA_L2PROMO") ? atoi(getenv("BTK_TMA_L2PROMO")) : -1;
  if (rows_env >= 8 && rows_env <= 256 && p->geo.D % rows_env == 0) rows = rows_env;
  const CUtensorMapL2promotion promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo_env == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  std::vector<CUtensorMap> maps(n > 0 ? n : 1);
  memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
  for (int i = 0; i < n; i++) {
    const long long off = p->sig[1 + 3 * i], T = p->sig[2 + 3 * i];
    if (T <= 0) continue;                                      // no work item refers to an empty recording
    const cuuint64_t gdim[2] = {(cuuint64_t)p->C, (cuuint64_t)T};
    const cuuint64_t gstr[1] = {(cuuint64_t)p->C * sizeof(float)};
    const cuuint32_t box[2] = {4u, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(d_pcm + off), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return nullptr;              // earlier launches may still read the old maps
  if (p->d_tmaps.reserve(maps.size() * sizeof(CUtensorMap)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(p->d_tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  p->tmap_sig = key;
  p->tma_rows = rows;
  return p->d_tmaps.p;
}

// Launch the work items [w0, w1) of the prepared batch.
static int chain_launch(btkb200_plan* p, const float* d_pcm, float* d_out, int w0, int w1, cudaStream_t st,
                        const cf* wts = nullptr, long long wts_stride = 0) {
  if (w1 <= w0) return BTKB200_OK;
  ChainParams c;
  c.pcm = d_pcm; c.out = d_out; c.recs = (const RecDesc*)p->d_recs.p; c.work = (const WorkItem*)p->d_work.p + w0;
  c.taps_h = p->d_taps_h; c.taps_g = p->d_taps_g; c.wts = wts ? wts : p->d_wts_chain; c.wts_stride = wts_stride;
  c.twa = p->d_twa; c.twb = p->d_twb;
  c.one_cta = p->one_cta;
  c.no_prefetch = p->no_prefetch;
  c.C = p->C; c.Cpad = p->Cpad; c.m = p->geo.m; c.pd_s = p->geo.pd_s; c.laN = p->geo.laN; c.gain = p->gain;
  c.cluster = p->use_ws ? p->cluster : 1;
  c.tmaps = p->use_ws ? chain_tensor_maps(p, d_pcm, st) : nullptr;
  c.tma_rows = p->tma_rows;
  c.item_begin = nullptr; c.item_q = 0; c.item0 = 0; c.n_items = 0; c.n_rec = p->sig.empty() ? 0 : (int)p->sig[0];
  c.cta_begin = nullptr; c.cta_n = 0;
  static const int syn_env = getenv("BTK_WS_SYN") ? atoi(getenv("BTK_WS_SYN")) : 1;   // A/B runs
  static const int dual_env = getenv("BTK_WS_DUAL") ? atoi(getenv("BTK_WS_DUAL")) : 1;   // A/B: two-channel windowing (bit 1)
  c.no_syn = (syn_env ? 0 : 1) | (dual_env ? 0 : 2);
  int n_cta = w1 - w0;
  if (p->use_ws && p->persist) {
    c.work = nullptr;
    c.item_begin = (const int*)p->d_work.p; c.item_q = p->item_q; c.item0 = w0; c.n_items = w1 - w0;
    n_cta = c.n_items;           // the launcher clamps the grid to the CTAs (clusters) resident at once
    if (p->cta_n > 0 && w0 == 0 && w1 == p->cached_n_work) {     // the whole batch: iteration-balanced CTA boundaries
      c.cta_begin = c.item_begin + (c.n_rec + 1);
      c.cta_n = p->cta_n;
      n_cta = p->cta_n;
    }
  }
  if (p->use_ws) CK(p, launch_chain_ws(p->geo.M, p->geo.R, c, n_cta, st));
  else CK(p, launch_chain(p->geo.M, p->geo.R, c, w1 - w0, st));
  p->launched_chain = true;
  p->launches++;
  return BTKB200_OK;
}

int btkb200_chain_batch_dev(btkb200_plan* p, const float* d_pcm, const long long* pcm_off, const long long* T,
                            const long long* out_off, int n, float* d_out, void* stream) {
  if (!p || !d_pcm || !d_out || !pcm_off || !T || !out_off || n < 0) return BTKB200_EINVAL;
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  if (!p->has_h || !p->has_g) return fail(p, BTKB200_ESTATE, "the fused chain needs both prototypes");
  if (n == 0) return BTKB200_OK;
  CK(p, cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = chain_prepare(p, pcm_off, T, out_off, n, st);
  if (rc) return rc;
  return chain_launch(p, d_pcm, d_out, 0, p->cached_n_work, st);
}

long btkb200_launch_count(const btkb200_plan* p) { return p ? p->launches : -1; }

int btkb200_plan_tune(btkb200_plan* p, int knob, int value) {
  if (!p) return BTKB200_EINVAL;
  const bool has_ws = chain_ws_frames_per_iter(p->geo.M, p->geo.R, p->geo.m) > 0;
  if (knob == BTKB200_TUNE_CHAIN_WS) {
    if (value > 0 && !has_ws)
      return fail(p, BTKB200_EUNSUPPORTED, "no warp-specialised chain kernel for M=%d r=%d m=%d", p->geo.M, p->geo.r, p->geo.m);
    p->tune_ws = value < 0 ? -1 : (value ? 1 : 0);
  } else if (knob == BTKB200_TUNE_CLUSTER) {
    if (value < 0 || value > 8) return fail(p, BTKB200_EUNSUPPORTED, "cluster size %d outside 0..8", value);
    if (value > 0 && (!has_ws || (p->Cpad / 4) % value != 0 || !chain_ws_cluster_ok(p->geo.M, p->geo.R, p->geo.m, value)))
      return fail(p, BTKB200_EUNSUPPORTED, "cluster of %d CTAs impossible for %d channels, M=%d", value, p->C, p->geo.M);
    p->tune_cluster = value;
  } else {
    return fail(p, BTKB200_EINVAL, "unknown tuning knob %d", knob);
  }
  p->sig.clear();      // the next launch rebuilds its work list
  return BTKB200_OK;
}

int btkb200_plan_tuning(const btkb200_plan* p, int knob) {
  if (!p || !p->launched_chain) return -1;
  if (knob == BTKB200_TUNE_CHAIN_WS) return p->use_ws;
  if (knob == BTKB200_TUNE_CLUSTER) return p->use_ws ? p->cluster : 1;
  return -1;
}

int btkb200_sync(btkb200_plan* p) {
  if (!p) return BTKB200_EINVAL;
  CK(p, cudaSetDevice(p->device));
  CK(p, cudaDeviceSynchronize());
  return BTKB200_OK;
}

// ------------------------------------------------------------------------------------------- host-buffer entry points
int btkb200_analysis(btkb200_plan* p, const float* pcm, long T, float* snap, long* n_frames) {
  if (!p || !pcm || !snap || T < 0) return BTKB200_EINVAL;
  CK(p, cudaSetDevice(p->device));
  const long F = p->geo.analysis_frames(T);
  const size_t bin = (size_t)T * p->C * sizeof(float), bout = (size_t)F * p->geo.B * p->C * sizeof(cf);
  CK(p, p->d_in.reserve(bin ? bin : 16));
  CK(p, p->d_out.reserve(bout ? bout : 16));
  if (bin) CK(p, cudaMemcpyAsync(p->d_in.p, pcm, bin, cudaMemcpyHostToDevice, p->stream));
  int rc = btkb200_analysis_dev(p, (const float*)p->d_in.p, T, (float*)p->d_out.p, p->stream);
  if (rc) return rc;
  if (bout) CK(p, cudaMemcpyAsync(snap, p->d_out.p, bout, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  if (n_frames) *n_frames = F;
  return BTKB200_OK;
}

int btkb200_beamform(btkb200_plan* p, const float* snap, long F, float* Y) {
  if (!p || !snap || !Y || F < 0) return BTKB200_EINVAL;
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  CK(p, cudaSetDevice(p->device));
  const size_t bin = (size_t)F * p->geo.B * p->C * sizeof(cf), bout = (size_t)F * p->geo.B * sizeof(cf);
  if (F == 0) return BTKB200_OK;
  CK(p, p->d_in.reserve(bin));
  CK(p, p->d_out.reserve(bout));
  CK(p, cudaMemcpyAsync(p->d_in.p, snap, bin, cudaMemcpyHostToDevice, p->stream));
  int rc = btkb200_beamform_dev(p, (const float*)p->d_in.p, F, (float*)p->d_out.p, p->stream);
  if (rc) return rc;
  CK(p, cudaMemcpyAsync(Y, p->d_out.p, bout, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_beamform_zelinski(btkb200_plan* p, const float* snap, long F, double alpha, int type, int min_frames, float* Y,
                              float* W) {
  if (!p || !snap || !Y || F < 0) return BTKB200_EINVAL;
  CK(p, cudaSetDevice(p->device));
  const size_t bin = (size_t)F * p->geo.B * p->C * sizeof(cf), bY = (size_t)F * p->geo.B * sizeof(cf);
  const size_t bW = (size_t)F * p->geo.B * sizeof(float);
  if (F == 0) return BTKB200_OK;
  CK(p, p->d_in.reserve(bin));
  CK(p, p->d_out.reserve(bY + bW + 16));
  float* dW = (float*)((char*)p->d_out.p + ((bY + 15) / 16) * 16);
  CK(p, cudaMemcpyAsync(p->d_in.p, snap, bin, cudaMemcpyHostToDevice, p->stream));
  int rc = btkb200_beamform_zelinski_dev(p, (const float*)p->d_in.p, F, alpha, type, min_frames, (float*)p->d_out.p, dW, p->stream);
  if (rc) return rc;
  CK(p, cudaMemcpyAsync(Y, p->d_out.p, bY, cudaMemcpyDeviceToHost, p->stream));
  if (W) CK(p, cudaMemcpyAsync(W, dW, bW, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_chain_zelinski(btkb200_plan* p, const float* pcm, long T, double alpha, int type, int min_frames, float* out) {
  if (!p || !pcm || !out || T < 0) return BTKB200_EINVAL;
  if (!p->has_h || !p->has_g) return fail(p, BTKB200_ESTATE, "the chain needs both prototypes");
  CK(p, cudaSetDevice(p->device));
  const long F = p->geo.analysis_frames(T), nout = p->geo.synthesis_frames((int)F);
  const size_t bin = (size_t)T * p->C * sizeof(float), bsnap = (size_t)F * p->geo.B * p->C * sizeof(cf);
  const size_t bY = (size_t)F * p->geo.B * sizeof(cf), bout = (size_t)nout * p->geo.D * sizeof(float);
  CK(p, p->d_in.reserve(bin ? bin : 16));
  CK(p, p->d_out.reserve(bsnap ? bsnap : 16));
  CK(p, p->d_aux2.reserve(bY + bout + 32));
  float* dY = (float*)p->d_aux2.p;
  float* dout = (float*)((char*)p->d_aux2.p + ((bY + 15) / 16) * 16);
  if (bin) CK(p, cudaMemcpyAsync(p->d_in.p, pcm, bin, cudaMemcpyHostToDevice, p->stream));
  int rc = btkb200_analysis_dev(p, (const float*)p->d_in.p, T, (float*)p->d_out.p, p->stream);
  if (rc) return rc;
  rc = btkb200_beamform_zelinski_dev(p, (const float*)p->d_out.p, F, alpha, type, min_frames, dY, nullptr, p->stream);
  if (rc) return rc;
  rc = btkb200_synthesis_dev(p, dY, F, dout, p->stream);
  if (rc) return rc;
  if (bout) CK(p, cudaMemcpyAsync(out, dout, bout, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

// Many recordings through analysis -> beamformer -> Zelinski post-filter -> synthesis: all descriptors are uploaded once,
// the uploads run ahead on the copy stream, the four kernels of a recording follow each other on the plan's stream and
// the downloads trail on a third stream.
int btkb200_chain_zelinski_batch(btkb200_plan* p, const float* const* pcm, const long* T, int n, double alpha, int type,
                                 int min_frames, float* const* out) {
  if (!p || !pcm || !T || !out || n < 0) return BTKB200_EINVAL;
  if (!p->has_h || !p->has_g) return fail(p, BTKB200_ESTATE, "the chain needs both prototypes");
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  if (!p->has_manifold) return fail(p, BTKB200_ESTATE, "set beamformer's weights");
  if (p->C <= 1) return fail(p, BTKB200_EINVAL, "The number of channels %d is <= 1", p->C);
  if (n == 0) return BTKB200_OK;
  CK(p, cudaSetDevice(p->device));
  const int B = p->geo.B, C = p->C, M = p->geo.M, D = p->geo.D, Cpad = p->Cpad;
  std::vector<long long> poff(n), ooff(n);
  long long pin = 0, pout = 0;
  for (int i = 0; i < n; i++) {
    if (T[i] < 0 || !pcm[i] || !out[i]) return fail(p, BTKB200_EINVAL, "recording %d: bad buffer or length", i);
    poff[i] = pin; ooff[i] = pout;
    pin += ((long long)T[i] * C + 3) / 4 * 4;
    pout += ((long long)p->geo.chain_frames(T[i]) * D + 3) / 4 * 4;
  }
  // descriptors of the analysis and synthesis launches of every recording
  std::vector<RecDesc> arecs(n), srecs(n);
  std::vector<WorkItem> awork, swork;
  std::vector<int> awb(n + 1, 0), swb(n + 1, 0), aslices(n, 1);
  const int Wa = fb_frames_per_iter(M, p->geo.R), H = p->geo.m * p->geo.R - 1, n_cg = Cpad / 4;
  size_t max_snap = 16, max_Y = 16;
  for (int i = 0; i < n; i++) {
    const long F = p->geo.analysis_frames(T[i]);
    const int nout = p->geo.synthesis_frames((int)F);
    arecs[i].pcm_off = poff[i]; arecs[i].out_off = 0; arecs[i].T = (int)T[i]; arecs[i].nblk = (int)F;
    srecs[i].pcm_off = 0; srecs[i].out_off = ooff[i]; srecs[i].T = (int)F; srecs[i].nblk = nout;
    std::vector<RecDesc> one(1, arecs[i]);
    std::vector<WorkItem> w1;
    long long chunk = ((long long)F + 148 * 4 - 1) / (148 * 4);
    chunk = (chunk + Wa - 1) / Wa * Wa;
    if (chunk < Wa) chunk = Wa;
    build_work(one, (int)chunk, w1);
    for (size_t k = 0; k < w1.size(); k++) { w1[k].rec = i; awork.push_back(w1[k]); }
    awb[i + 1] = (int)awork.size();
    int slices = w1.empty() ? 1 : (int)(2 * 148 / w1.size());
    if (slices < 1) slices = 1;
    if (slices > n_cg) slices = n_cg;
    aslices[i] = slices;
    one[0] = srecs[i];
    w1.clear();
    if (nout > 0) build_work(one, choose_chunk(nout, H, Wa), w1);
    for (size_t k = 0; k < w1.size(); k++) { w1[k].rec = i; swork.push_back(w1[k]); }
    swb[i + 1] = (int)swork.size();
    if ((size_t)F * B * C * sizeof(cf) > max_snap) max_snap = (size_t)F * B * C * sizeof(cf);
    if ((size_t)F * B * sizeof(cf) > max_Y) max_Y = (size_t)F * B * sizeof(cf);
  }
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_arecs = 0, o_srecs = o_arecs + al(n * sizeof(RecDesc));
  const size_t o_awork = o_srecs + al(n * sizeof(RecDesc));
  const size_t o_swork = o_awork + al((awork.size() ? awork.size() : 1) * sizeof(WorkItem));
  const size_t o_Y = o_swork + al((swork.size() ? swork.size() : 1) * sizeof(WorkItem));
  const size_t o_stat = o_Y + al(max_Y);
  const size_t o_snap = o_stat + al(max_Y * 2 + (max_Y / 8 / 32 + B + 64) * 32);   // float4 per (frame, bin) + scan segments
  CK(p, p->d_adapt.reserve(o_snap + al(max_snap)));
  CK(p, p->d_in.reserve((size_t)(pin ? pin : 4) * sizeof(float)));
  CK(p, p->d_out.reserve((size_t)(pout ? pout : 4) * sizeof(float)));
  if (!p->s_in) CK(p, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  if (!p->s_out) CK(p, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  char* base = (char*)p->d_adapt.p;
  CK(p, cudaStreamSynchronize(p->stream));
  CK(p, cudaMemcpy(base + o_arecs, arecs.data(), n * sizeof(RecDesc), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(base + o_srecs, srecs.data(), n * sizeof(RecDesc), cudaMemcpyHostToDevice));
  if (!awork.empty()) CK(p, cudaMemcpy(base + o_awork, awork.data(), awork.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
  if (!swork.empty()) CK(p, cudaMemcpy(base + o_swork, swork.data(), swork.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
  cf* dY = (cf*)(base + o_Y);
  float4* dstat = (float4*)(base + o_stat);
  cf* dsnap = (cf*)(base + o_snap);
  while ((int)p->ev_in.size() < n) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_in.push_back(e); }
  while ((int)p->ev_k.size() < n) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_k.push_back(e); }
  for (int i = 0; i < n; i++) {
    if (T[i] > 0)
      CK(p, cudaMemcpyAsync((float*)p->d_in.p + poff[i], pcm[i], (size_t)T[i] * C * sizeof(float), cudaMemcpyHostToDevice, p->s_in));
    CK(p, cudaEventRecord(p->ev_in[i], p->s_in));
  }
  for (int i = 0; i < n; i++) {
    CK(p, cudaStreamWaitEvent(p->stream, p->ev_in[i], 0));
    const long F = arecs[i].nblk;
    if (awb[i + 1] > awb[i]) {
      AnalysisParams a;
      a.pcm = (const float*)p->d_in.p; a.snap = dsnap; a.recs = (const RecDesc*)(base + o_arecs);
      a.work = (const WorkItem*)(base + o_awork) + awb[i];
      a.taps_h = p->d_taps_h; a.twa = p->d_twa; a.twb = p->d_twb; a.C = C; a.Cpad = Cpad; a.m = p->geo.m; a.laN = p->geo.laN;
      a.cg_slices = aslices[i];
      CK(p, launch_analysis(M, p->geo.R, a, (awb[i + 1] - awb[i]) * aslices[i], p->stream));
      CK(p, launch_beamform_zelinski(dsnap, p->d_w, p->d_ta, dY, dstat, nullptr, F, B, C, alpha, type, min_frames, p->stream));
      p->launches += 5;
    }
    if (swb[i + 1] > swb[i]) {
      SynthesisParams s;
      s.Y = dY; s.out = (float*)p->d_out.p; s.recs = (const RecDesc*)(base + o_srecs); s.work = (const WorkItem*)(base + o_swork) + swb[i];
      s.taps_g = p->d_taps_g; s.twa = p->d_twa; s.twb = p->d_twb; s.m = p->geo.m; s.pd_s = p->geo.pd_s; s.gain = p->gain;
      CK(p, launch_synthesis(M, p->geo.R, s, swb[i + 1] - swb[i], p->stream));
      p->launches++;
    }
    CK(p, cudaEventRecord(p->ev_k[i], p->stream));
    CK(p, cudaStreamWaitEvent(p->s_out, p->ev_k[i], 0));
    const size_t b = (size_t)p->geo.chain_frames(T[i]) * D * sizeof(float);
    if (b) CK(p, cudaMemcpyAsync(out[i], (float*)p->d_out.p + ooff[i], b, cudaMemcpyDeviceToHost, p->s_out));
  }
  CK(p, cudaStreamSynchronize(p->s_out));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_synthesis(btkb200_plan* p, const float* Y, long F, float* out, long* n_out_frames) {
  if (!p || !Y || !out || F < 0) return BTKB200_EINVAL;
  CK(p, cudaSetDevice(p->device));
  const long nout = p->geo.synthesis_frames((int)F);
  if (n_out_frames) *n_out_frames = nout;
  if (nout == 0) return p->has_g ? BTKB200_OK : fail(p, BTKB200_ESTATE, "plan was created without a synthesis prototype");
  const size_t bin = (size_t)F * p->geo.B * sizeof(cf), bout = (size_t)nout * p->geo.D * sizeof(float);
  CK(p, p->d_in.reserve(bin));
  CK(p, p->d_out.reserve(bout));
  CK(p, cudaMemcpyAsync(p->d_in.p, Y, bin, cudaMemcpyHostToDevice, p->stream));
  int rc = btkb200_synthesis_dev(p, (const float*)p->d_in.p, F, (float*)p->d_out.p, p->stream);
  if (rc) return rc;
  CK(p, cudaMemcpyAsync(out, p->d_out.p, bout, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_covariance(btkb200_plan* p, const float* snap, long F, const double* frame_weights, int conjugate, double* R) {
  if (!p || !snap || !frame_weights || !R || F < 0) return BTKB200_EINVAL;
  if (p->C > 64) return fail(p, BTKB200_EUNSUPPORTED, "covariance supports at most 64 channels (got %d)", p->C);
  // the tensor-core kernel scales the rows by sqrt(weight): a negative or non-finite weight has no such factor (the
  // recursions of the reference only ever produce weights in [0, 1])
  for (long f = 0; f < F; f++)
    if (!(frame_weights[f] >= 0.0) || !std::isfinite(frame_weights[f]))
      return fail(p, BTKB200_EINVAL, "frame weight %ld is negative or not finite (%g)", f, frame_weights[f]);
  CK(p, cudaSetDevice(p->device));
  const int B = p->geo.B, C = p->C;
  const size_t bsnap = (size_t)F * B * C * sizeof(cf), bw = (size_t)F * sizeof(double), bR = (size_t)B * C * C * sizeof(double2);
  CK(p, p->d_in.reserve(bsnap ? bsnap : 16));
  CK(p, p->d_aux2.reserve(bw + 16 + bR));
  double* dwt = (double*)p->d_aux2.p;
  double2* dR = (double2*)((char*)p->d_aux2.p + ((bw + 15) / 16) * 16);
  CK(p, cudaMemsetAsync(dR, 0, bR, p->stream));
  if (F > 0) {
    CK(p, cudaMemcpyAsync(p->d_in.p, snap, bsnap, cudaMemcpyHostToDevice, p->stream));
    CK(p, cudaMemcpyAsync(dwt, frame_weights, bw, cudaMemcpyHostToDevice, p->stream));
    CK(p, launch_covariance((const cf*)p->d_in.p, dwt, dR, F, B, C, conjugate ? 1 : 0, p->stream));
    p->launches++;
  }
  CK(p, cudaMemcpyAsync(R, dR, bR, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_estimate_covariance(btkb200_plan* p, const float* pcm, long T, double forget, long last_frame, int conjugate) {
  if (!p || !pcm || T < 0 || !(forget >= 0.0 && forget <= 1.0)) return BTKB200_EINVAL;
  if (!p->has_h) return fail(p, BTKB200_ESTATE, "plan was created without an analysis prototype");
  if (p->C > 64) return fail(p, BTKB200_EUNSUPPORTED, "covariance supports at most 64 channels (got %d)", p->C);
  CK(p, cudaSetDevice(p->device));
  const int B = p->geo.B, C = p->C;
  const long F = p->geo.analysis_frames(T);
  long Fu = (last_frame >= 0 && last_frame + 1 < F) ? last_frame + 1 : F;      // frames that adapt
  // frame weights of the recursion unrolled: frame f contributes (1-ff) ff^(Fu-1-f); the Python flavour starts from
  // S = x0 x0^H (weight ff^(Fu-1) for frame 0), the C++ flavour from R = 0
  std::vector<double> wt((size_t)(Fu > 0 ? Fu : 1), 0.0);
  double acc = 1.0;
  for (long f = Fu - 1; f >= 0; f--) { wt[f] = (1.0 - forget) * acc; acc *= forget; }
  if (conjugate && Fu > 0) wt[0] = pow(forget, (double)(Fu - 1));   // pow(0, 0) == 1: a single frame keeps S = x0 x0^H
  const size_t bin = (size_t)T * C * sizeof(float), bsnap = (size_t)F * B * C * sizeof(cf);
  const size_t bw = (size_t)wt.size() * sizeof(double), bR = (size_t)B * C * C * sizeof(double2);
  CK(p, p->d_in.reserve(bin ? bin : 16));
  CK(p, p->d_out.reserve(bsnap ? bsnap : 16));
  CK(p, p->d_aux2.reserve(bw + 16 + bR));
  double* dwt = (double*)p->d_aux2.p;
  double2* dR = (double2*)((char*)p->d_aux2.p + ((bw + 15) / 16) * 16);
  if (bin) CK(p, cudaMemcpyAsync(p->d_in.p, pcm, bin, cudaMemcpyHostToDevice, p->stream));
  CK(p, cudaMemsetAsync(dR, 0, bR, p->stream));
  int rc = btkb200_analysis_dev(p, (const float*)p->d_in.p, T, (float*)p->d_out.p, p->stream);
  if (rc) return rc;
  if (Fu > 0) {
    CK(p, cudaMemcpyAsync(dwt, wt.data(), bw, cudaMemcpyHostToDevice, p->stream));
    CK(p, launch_covariance((const cf*)p->d_out.p, dwt, dR, Fu, B, C, conjugate ? 1 : 0, p->stream));
    p->launches++;
  }
  CK(p, cudaMemcpyAsync(p->Rn.data(), dR, bR, cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  std::fill(p->Rn_set.begin(), p->Rn_set.end(), 1);
  return BTKB200_OK;
}

static int pcm_bytes_per_sample(int format) {
  return format == BTKB200_PCM_F32 ? 4 : format == BTKB200_PCM_S16 ? 2 : format == BTKB200_PCM_S24BE ? 3 : 0;
}

int btkb200_convert_pcm(btkb200_plan* p, const void* src, int format, long n, float* dst) {
  if (!p || !src || !dst || n < 0) return BTKB200_EINVAL;
  const int bpe = pcm_bytes_per_sample(format);
  if (!bpe) return fail(p, BTKB200_EINVAL, "unknown PCM format %d", format);
  if (n == 0) return BTKB200_OK;
  CK(p, cudaSetDevice(p->device));
  CK(p, p->d_in.reserve((size_t)n * sizeof(float)));
  if (format == BTKB200_PCM_F32) {
    CK(p, cudaMemcpyAsync(p->d_in.p, src, (size_t)n * 4, cudaMemcpyHostToDevice, p->stream));
  } else {
    CK(p, p->d_raw.reserve((size_t)n * bpe + 16));
    CK(p, cudaMemcpyAsync(p->d_raw.p, src, (size_t)n * bpe, cudaMemcpyHostToDevice, p->stream));
    CK(p, launch_ingest(format, p->d_raw.p, (float*)p->d_in.p, n, p->stream));
    p->launches++;
  }
  CK(p, cudaMemcpyAsync(dst, p->d_in.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

int btkb200_chain_batch(btkb200_plan* p, const float* const* pcm, const long* T, int n, float* const* out) {
  return btkb200_chain_batch_pcm(p, (const void* const*)pcm, BTKB200_PCM_F32, T, n, out);
}

int btkb200_chain_batch_pcm(btkb200_plan* p, const void* const* pcm, int format, const long* T, int n, float* const* out) {
  if (!p || !pcm || !T || !out || n < 0) return BTKB200_EINVAL;
  const int bpe = pcm_bytes_per_sample(format);
  if (!bpe) return fail(p, BTKB200_EINVAL, "unknown PCM format %d", format);
  if (!p->has_weights) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  if (!p->has_h || !p->has_g) return fail(p, BTKB200_ESTATE, "the fused chain needs both prototypes");
  if (n == 0) return BTKB200_OK;
  CK(p, cudaSetDevice(p->device));
  std::vector<long long> poff(n), ooff(n), Tl(n);
  long long pin = 0, pout = 0;
  for (int i = 0; i < n; i++) {
    if (T[i] < 0 || !pcm[i] || !out[i]) return fail(p, BTKB200_EINVAL, "recording %d: bad buffer or length", i);
    poff[i] = pin; ooff[i] = pout; Tl[i] = T[i];
    pin += ((long long)T[i] * p->C + 3) / 4 * 4;            // keep every recording 16-byte aligned for float4 loads
    pout += ((long long)p->geo.chain_frames(T[i]) * p->geo.D + 3) / 4 * 4;
  }
  CK(p, p->d_in.reserve((size_t)(pin ? pin : 4) * sizeof(float)));
  CK(p, p->d_out.reserve((size_t)(pout ? pout : 4) * sizeof(float)));
  if (format != BTKB200_PCM_F32) CK(p, p->d_raw.reserve((size_t)(pin ? pin : 4) * bpe + 16));
  if (!p->s_in) CK(p, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  if (!p->s_out) CK(p, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  int rc = chain_prepare(p, poff.data(), Tl.data(), ooff.data(), n, p->stream);
  if (rc) return rc;
  // Three-stage pipeline over groups of recordings: H2D on s_in, kernel on the plan's stream, D2H on s_out, so the
  // upload of group k+1, the transform of group k and the download of group k-1 overlap (PCIe is full duplex).
  // Group boundaries balance input bytes; a group's launch covers exactly the work items of its recordings.
  const int G = n < 8 ? n : 8;
  while ((int)p->ev_in.size() < G) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_in.push_back(e); }
  while ((int)p->ev_k.size() < G) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_k.push_back(e); }
  int r0 = 0;
  for (int g = 0; g < G; g++) {
    const long long want = pin * (g + 1) / G;
    int r1 = r0;
    while (r1 < n && (r1 == r0 || poff[r1] < want)) r1++;
    if (g == G - 1) r1 = n;
    // raw formats land in d_raw at the same ELEMENT offsets (multiples of 4 elements: 8- / 12-byte aligned) and are
    // converted to float32 in d_in by one element-wise kernel per group, ahead of the group's chain launch
    char* dst_base = format == BTKB200_PCM_F32 ? (char*)p->d_in.p : (char*)p->d_raw.p;
    for (int i = r0; i < r1; i++)
      if (T[i] > 0)
        CK(p, cudaMemcpyAsync(dst_base + (size_t)poff[i] * bpe, pcm[i], (size_t)T[i] * p->C * bpe, cudaMemcpyHostToDevice, p->s_in));
    CK(p, cudaEventRecord(p->ev_in[g], p->s_in));
    CK(p, cudaStreamWaitEvent(p->stream, p->ev_in[g], 0));
    if (format != BTKB200_PCM_F32) {
      const long long e0 = poff[r0], e1 = r1 < n ? poff[r1] : pin;
      CK(p, launch_ingest(format, (const char*)p->d_raw.p + (size_t)e0 * bpe, (float*)p->d_in.p + e0, e1 - e0, p->stream));
      p->launches++;
    }
    rc = chain_launch(p, (const float*)p->d_in.p, (float*)p->d_out.p, p->rec_work_begin[r0], p->rec_work_begin[r1], p->stream);
    if (rc) return rc;
    CK(p, cudaEventRecord(p->ev_k[g], p->stream));
    CK(p, cudaStreamWaitEvent(p->s_out, p->ev_k[g], 0));
    for (int i = r0; i < r1; i++) {
      const size_t b = (size_t)p->geo.chain_frames(T[i]) * p->geo.D * sizeof(float);
      if (b) CK(p, cudaMemcpyAsync(out[i], (float*)p->d_out.p + ooff[i], b, cudaMemcpyDeviceToHost, p->s_out));
    }
    r0 = r1;
    if (r0 >= n) break;
  }
  CK(p, cudaStreamSynchronize(p->s_out));
  CK(p, cudaStreamSynchronize(p->stream));
  return BTKB200_OK;
}

// ------------------------------------------------------------------------------------------- adaptive MVDR, batched
static void recursion_weights(std::vector<double>& wt, long Fu, double forget, int conjugate) {
  // frame f of the unrolled recursion contributes (1-ff) ff^(Fu-1-f); the Python flavour starts from S = x0 x0^H
  wt.assign((size_t)(Fu > 0 ? Fu : 1), 0.0);
  double acc = 1.0;
  for (long f = Fu - 1; f >= 0; f--) { wt[f] = (1.0 - forget) * acc; acc *= forget; }
  if (conjugate && Fu > 0) wt[0] = pow(forget, (double)(Fu - 1));
}

int btkb200_mvdr_chain_batch(btkb200_plan* p, const float* const* pcm, const long* T, int n, const btkb200_mvdr_adapt* cfg,
                             float* const* out, int* n_fallback) {
  if (!p || !pcm || !T || !out || !cfg || n < 0) return BTKB200_EINVAL;
  if (!(cfg->forget >= 0.0 && cfg->forget <= 1.0)) return fail(p, BTKB200_EINVAL, "forgetting factor %g outside [0, 1]", cfg->forget);
  if (!p->has_h || !p->has_g) return fail(p, BTKB200_ESTATE, "the fused chain needs both prototypes");
  if (!p->has_manifold) return fail(p, BTKB200_ESTATE, "call calcArrayManifoldVectorsX() once");
  if (p->C > 64) return fail(p, BTKB200_EUNSUPPORTED, "MVDR adaptation supports at most 64 channels (got %d)", p->C);
  if (n == 0) return BTKB200_OK;
  CK(p, cudaSetDevice(p->device));
  const int B = p->geo.B, C = p->C, M = p->geo.M, D = p->geo.D, Cpad = p->Cpad;
  // ---- batch layout (like btkb200_chain_batch)
  std::vector<long long> poff(n), ooff(n), Tl(n);
  long long pin = 0, pout = 0;
  for (int i = 0; i < n; i++) {
    if (T[i] < 0 || !pcm[i] || !out[i]) return fail(p, BTKB200_EINVAL, "recording %d: bad buffer or length", i);
    poff[i] = pin; ooff[i] = pout; Tl[i] = T[i];
    pin += ((long long)T[i] * C + 3) / 4 * 4;
    pout += ((long long)p->geo.chain_frames(T[i]) * D + 3) / 4 * 4;
  }
  CK(p, p->d_in.reserve((size_t)(pin ? pin : 4) * sizeof(float)));
  CK(p, p->d_out.reserve((size_t)(pout ? pout : 4) * sizeof(float)));
  if (!p->s_in) CK(p, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  int rc = chain_prepare(p, poff.data(), Tl.data(), ooff.data(), n, p->stream);
  if (rc) return rc;
  // ---- Adaptation runs group by group: ONE launch per stage over all recordings of a group (analysis of the adapting
  // lead-ins, tensor-core covariance with the recording as grid.z, diagonal loading over n B bins, n B solve CTAs, n weight
  // tables), then the fused-chain launch of that group with per-recording weight tables and the download of its outputs.
  // The first sessions' version issued the five small launches recording by recording: at 64 channels the latency-bound
  // solve alone (140 us, 257 CTAs) made 1 024 utterances cost more than their chain.
  std::vector<long> Fu(n);
  std::vector<long long> fwoff(n);
  std::vector<double> fw_all, wt;
  const int Wa = fb_frames_per_iter(M, p->geo.R), n_cg = Cpad / 4;
  long Fa_max = 0;
  for (int i = 0; i < n; i++) {
    const long F = p->geo.analysis_frames(T[i]);
    Fu[i] = (cfg->last_frame >= 0 && cfg->last_frame + 1 < F) ? cfg->last_frame + 1 : F;
    if (Fu[i] > Fa_max) Fa_max = Fu[i];
    recursion_weights(wt, Fu[i], cfg->forget, cfg->conjugate);
    fwoff[i] = (long long)fw_all.size();
    fw_all.insert(fw_all.end(), wt.begin(), wt.end());
  }
  const size_t snap_stride = (size_t)(Fa_max > 0 ? Fa_max : 1) * B * C;           // complex elements per recording
  const size_t per_rec = snap_stride * sizeof(cf) + (size_t)B * C * C * sizeof(double2) + (size_t)B * C * sizeof(double2);
  // recordings per group: at most 3 GB of snapshots + matrices in flight
  const size_t budget = (size_t)3 << 30;
  int Gn = per_rec > 0 && budget / per_rec < 4096 ? (int)(budget / per_rec) : 4096;
  // at least eight groups for a batch of 16 or more: the upload of group k+1 and the download of group k-1 overlap the
  // kernels of group k (one group = upload everything, then compute, then download: cfg3 +17 % end to end)
  if (n >= 16 && Gn > (n + 7) / 8) Gn = (n + 7) / 8;
  if (Gn < 1) Gn = 1;
  if (Gn > n) Gn = n;
  if (Gn > 4096) Gn = 4096;
  std::vector<RecDesc> arecs(n);
  std::vector<WorkItem> awork;
  std::vector<int> awb(n + 1, 0);
  std::vector<CovRec> crecs(n);
  for (int i = 0; i < n; i++) {
    // frames 0 .. Fu-1 only look at samples below (Fu + laN) D: analyse just that lead-in
    long Ta = (long)((long long)(Fu[i] + p->geo.laN) * D);
    if (Ta > T[i]) Ta = T[i];
    arecs[i].pcm_off = poff[i]; arecs[i].out_off = (long long)(i % Gn) * (long long)snap_stride; arecs[i].T = (int)Ta; arecs[i].nblk = (int)Fu[i];
    std::vector<RecDesc> one(1, arecs[i]);
    std::vector<WorkItem> w1;
    long long chunk = ((long long)Fu[i] + 4 - 1) / 4;          // about four CTAs per lead-in: a group supplies the parallelism
    chunk = (chunk + Wa - 1) / Wa * Wa;
    if (chunk < Wa) chunk = Wa;
    build_work(one, (int)chunk, w1);
    for (size_t k = 0; k < w1.size(); k++) { w1[k].rec = i; awork.push_back(w1[k]); }
    awb[i + 1] = (int)awork.size();
    crecs[i].snap_off = (long long)(i % Gn) * (long long)snap_stride; crecs[i].wt_off = fwoff[i]; crecs[i].F = Fu[i];
  }
  std::vector<int> binmap;
  if (!build_weight_binmap(M, binmap)) return fail(p, BTKB200_EUNSUPPORTED, "no weight layout for M=%d", M);
  // ---- one scratch allocation: descriptors | recursion weights | bin map | manifold | fallback flags | tables (all n) |
  //      R, w, snapshots (one group)
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_arecs = 0, o_awork = o_arecs + al(arecs.size() * sizeof(RecDesc));
  const size_t o_crecs = o_awork + al((awork.size() ? awork.size() : 1) * sizeof(WorkItem));
  const size_t o_fw = o_crecs + al(crecs.size() * sizeof(CovRec));
  const size_t o_map = o_fw + al(fw_all.size() * sizeof(double));
  const size_t o_d = o_map + al(binmap.size() * sizeof(int));
  const size_t o_fb = o_d + al((size_t)B * C * sizeof(double2));
  const size_t o_tab = o_fb + al((size_t)n * B * sizeof(int));
  const size_t o_R = o_tab + al((size_t)n * Cpad * M * sizeof(cf));
  const size_t o_w = o_R + al((size_t)Gn * B * C * C * sizeof(double2));
  const size_t o_snap = o_w + al((size_t)Gn * B * C * sizeof(double2));
  CK(p, p->d_adapt.reserve(o_snap + al((size_t)Gn * snap_stride * sizeof(cf))));
  char* base = (char*)p->d_adapt.p;
  CK(p, cudaStreamSynchronize(p->stream));
  CK(p, cudaMemcpy(base + o_arecs, arecs.data(), arecs.size() * sizeof(RecDesc), cudaMemcpyHostToDevice));
  if (!awork.empty()) CK(p, cudaMemcpy(base + o_awork, awork.data(), awork.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(base + o_crecs, crecs.data(), crecs.size() * sizeof(CovRec), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(base + o_fw, fw_all.data(), fw_all.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(base + o_map, binmap.data(), binmap.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(p, cudaMemcpy(base + o_d, p->wq.data(), (size_t)B * C * sizeof(double2), cudaMemcpyHostToDevice));
  double2* dR = (double2*)(base + o_R);
  double2* dw = (double2*)(base + o_w);
  int* dfb = (int*)(base + o_fb);
  cf* dtab = (cf*)(base + o_tab);
  cf* dsnap = (cf*)(base + o_snap);
  if (!p->s_out) CK(p, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  const int n_groups = (n + Gn - 1) / Gn;
  while ((int)p->ev_in.size() < n_groups) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_in.push_back(e); }
  while ((int)p->ev_k.size() < n_groups) { cudaEvent_t e; CK(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->ev_k.push_back(e); }
  // the uploads of all groups run ahead on the copy stream
  for (int g = 0; g < n_groups; g++) {
    const int r0 = g * Gn, r1 = r0 + Gn < n ? r0 + Gn : n;
    for (int i = r0; i < r1; i++)
      if (T[i] > 0)
        CK(p, cudaMemcpyAsync((float*)p->d_in.p + poff[i], pcm[i], (size_t)T[i] * C * sizeof(float), cudaMemcpyHostToDevice, p->s_in));
    CK(p, cudaEventRecord(p->ev_in[g], p->s_in));
  }
  static const bool cov_simt = getenv("BTK_COV_SIMT") && getenv("BTK_COV_SIMT")[0] == '1';
  for (int g = 0; g < n_groups; g++) {
    const int r0 = g * Gn, r1 = r0 + Gn < n ? r0 + Gn : n, ng = r1 - r0;
    CK(p, cudaStreamWaitEvent(p->stream, p->ev_in[g], 0));
    CK(p, cudaMemsetAsync(dR, 0, (size_t)ng * B * C * C * sizeof(double2), p->stream));
    const int nw = awb[r1] - awb[r0];
    long Fg = 0;
    for (int i = r0; i < r1; i++) if (Fu[i] > Fg) Fg = Fu[i];
    if (nw > 0 && Fg > 0) {
      AnalysisParams a;
      a.pcm = (const float*)p->d_in.p; a.snap = dsnap; a.recs = (const RecDesc*)(base + o_arecs);
      a.work = (const WorkItem*)(base + o_awork) + awb[r0];
      a.taps_h = p->d_taps_h; a.twa = p->d_twa; a.twb = p->d_twb; a.C = C; a.Cpad = Cpad; a.m = p->geo.m; a.laN = p->geo.laN;
      int slices = (int)(2 * 148 / nw);                 // few work items (a small batch): spread the channel groups over CTAs
      if (slices < 1) slices = 1;
      if (slices > n_cg) slices = n_cg;
      a.cg_slices = slices;
      CK(p, launch_analysis(M, p->geo.R, a, nw * slices, p->stream));
      p->launches++;
      if (cov_simt) {
        for (int i = r0; i < r1; i++)
          if (Fu[i] > 0)
            CK(p, launch_covariance(dsnap + crecs[i].snap_off, (const double*)(base + o_fw) + fwoff[i], dR + (size_t)(i - r0) * B * C * C,
                                    Fu[i], B, C, cfg->conjugate ? 1 : 0, p->stream));
        p->launches += ng;
      } else {
        CK(p, launch_covariance_tc_batch(dsnap, (const double*)(base + o_fw), dR, (const CovRec*)(base + o_crecs) + r0, ng, Fg, B, C,
                                         cfg->conjugate ? 1 : 0, p->stream));
        p->launches++;
      }
    }
    if (cfg->load_abs != 0.0 || cfg->load_rel != 0.0) {
      CK(p, launch_diag_load(dR, ng * B, C, (float)cfg->load_abs, cfg->load_rel, p->stream));
      p->launches++;
    }
    // x x^H covariances (+ loading) are Hermitian: the L D L^H kernel; the complex-symmetric x x^T flavour of
    // SpectralMatrixArray::update (conjugate = 0) keeps the general pivoted elimination.  BTK_MVDR_CHOL=0: A/B runs.
    static const int chol_env = getenv("BTK_MVDR_CHOL") ? atoi(getenv("BTK_MVDR_CHOL")) : 1;
    if (cfg->conjugate && chol_env)
      CK(p, launch_mvdr_chol(dR, (const double2*)(base + o_d), dw, dfb + (size_t)r0 * B, B, C, cfg->dThreshold, p->stream, ng));
    else
      CK(p, launch_mvdr_solve(dR, (const double2*)(base + o_d), dw, dfb + (size_t)r0 * B, B, C, cfg->dThreshold, p->stream, ng));
    CK(p, launch_weight_table(dw, (const int*)(base + o_map), dtab + (size_t)r0 * Cpad * M, M, C, Cpad, p->stream, ng));
    p->launches += 2;
    // the fused chain of this group, every recording with its own weight table, then its outputs go home
    rc = chain_launch(p, (const float*)p->d_in.p, (float*)p->d_out.p, p->rec_work_begin[r0], p->rec_work_begin[r1], p->stream, dtab,
                      (long long)Cpad * M);
    if (rc) return rc;
    CK(p, cudaEventRecord(p->ev_k[g], p->stream));
    CK(p, cudaStreamWaitEvent(p->s_out, p->ev_k[g], 0));
    for (int i = r0; i < r1; i++) {
      const size_t b = (size_t)p->geo.chain_frames(T[i]) * D * sizeof(float);
      if (b) CK(p, cudaMemcpyAsync(out[i], (float*)p->d_out.p + ooff[i], b, cudaMemcpyDeviceToHost, p->s_out));
    }
  }
  std::vector<int> fb;
  if (n_fallback) { fb.resize((size_t)n * B); CK(p, cudaMemcpyAsync(fb.data(), dfb, fb.size() * sizeof(int), cudaMemcpyDeviceToHost, p->stream)); }
  CK(p, cudaStreamSynchronize(p->s_out));
  CK(p, cudaStreamSynchronize(p->stream));
  if (n_fallback)
    for (int i = 0; i < n; i++) { int sfb = 0; for (int b = 0; b < B; b++) sfb += fb[(size_t)i * B + b]; n_fallback[i] = sfb; }
  return BTKB200_OK;
}

int btkb200_chain(btkb200_plan* p, const float* pcm, long T, float* out) {
  const float* pp[1] = {pcm};
  float* oo[1] = {out};
  long TT[1] = {T};
  return btkb200_chain_batch(p, pp, TT, 1, oo);
}

int btkb200_chain_batch_multi(btkb200_plan* const* plans, int n_plans, const float* const* pcm, const long* T, int n,
                              float* const* out) {
  if (!plans || n_plans <= 0 || !pcm || !T || !out || n < 0) return BTKB200_EINVAL;
  // round-robin partition; every device's share is enqueued before any is waited for
  std::vector<std::vector<const float*> > sp(n_plans);
  std::vector<std::vector<float*> > so(n_plans);
  std::vector<std::vector<long> > sT(n_plans);
  for (int i = 0; i < n; i++) { sp[i % n_plans].push_back(pcm[i]); so[i % n_plans].push_back(out[i]); sT[i % n_plans].push_back(T[i]); }
  // Host-buffer entry points synchronise their own stream at the end, so issue the shares from
  // one host thread per device.
  int rc_all = BTKB200_OK;
#pragma omp parallel for num_threads(n_plans) if (n_plans > 1)
  for (int d = 0; d < n_plans; d++) {
    if (sp[d].empty()) continue;
    int rc = btkb200_chain_batch(plans[d], sp[d].data(), sT[d].data(), (int)sp[d].size(), so[d].data());
    if (rc != BTKB200_OK) {
#pragma omp critical
      rc_all = rc;
    }
  }
  return rc_all;
}

void* btkb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}

void btkb200_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
