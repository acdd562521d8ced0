// tests/emu/emu_chain.cc -- TEST HARNESS: runs the library's own tile programs (csrc/*.cuh, the exact code
// the sm_100a kernels execute) sequentially on the host, one CTA at a time, so that the CPU-only test
// tier can check index arithmetic against the oracle without a GPU.  Not part of the product and never
// loaded by it; there is no CPU fallback in the library.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "host_tables.h"
#include "staged_tiles.cuh"
#include "chain_ws.cuh"

using namespace btk;

#ifndef BTK_EMU_CASES
#define BTK_EMU_CASES(X) X(64, 1) X(64, 8) X(64, 2) X(128, 2) X(128, 4) X(256, 1) X(256, 2) X(256, 4) X(512, 1) X(512, 2) X(512, 4) X(512, 8) X(1024, 2) X(1024, 4)
#endif

template <int M, int PP = 1> struct HostCtx {
  std::vector<ChainThreadState<M, PP> > ts;
  int nt;
  explicit HostCtx(int n) : ts(n), nt(n) {}
  template <class F> void par(F f) { for (int t = 0; t < nt; t++) f(t, ts[t]); }
  void sync() {}
  void syncwarp() {}
};

template <int M, int R, int MT, int PP>
static int run_chain(int m, int dct, int C, int n_rec, const long long* Ts, const float* pcm, const long long* pcm_off,
                     float* out, const long long* out_off, const double* h, const double* g, const double* w_re_im,
                     int gain, int chunk) {
  typedef ChainCfg<M, R, MT, PP> K;
  BankGeom geo(M, m, /*r=*/0, dct);
  geo.r = 0; for (int x = R; x > 1; x >>= 1) geo.r++;
  geo = BankGeom(M, m, geo.r, dct);
  const int Cpad = (C + K::CG - 1) / K::CG * K::CG;
  std::vector<cf> twa, twb, gam;
  std::vector<float> gp, hf;
  build_fft_tables(M, twa, twb);
  build_synthesis_taps(g, M, m, gp);
  build_analysis_taps(h, M, m, R, hf);
  std::vector<zd> w((size_t)geo.B * C);
  for (size_t i = 0; i < w.size(); i++) w[i] = zd(w_re_im[2 * i], w_re_im[2 * i + 1]);
  build_chain_weight_table(w.data(), M, C, Cpad, gam);
  std::vector<RecDesc> recs(n_rec);
  for (int r = 0; r < n_rec; r++) {
    recs[r].pcm_off = pcm_off[r]; recs[r].out_off = out_off[r]; recs[r].T = (int)Ts[r]; recs[r].nblk = geo.chain_frames(Ts[r]);
  }
  std::vector<WorkItem> work;
  build_work(recs, chunk, work);
  ChainParams p;
  p.pcm = pcm; p.out = out; p.recs = recs.data(); p.work = work.data();
  p.taps_h = hf.data(); p.taps_g = gp.data(); p.wts = gam.data(); p.wts_stride = 0; p.one_cta = 0; p.no_prefetch = 0; p.twa = twa.data(); p.twb = twb.data();
  p.C = C; p.Cpad = Cpad; p.m = m; p.pd_s = geo.pd_s; p.laN = geo.laN; p.gain = gain;
  const ChainSmem L = chain_smem_layout<M, R, PP>(m);
  std::vector<unsigned char> smem(L.total + 64, 0xA5);   // poison: every read must have been written
  for (size_t wi = 0; wi < work.size(); wi++) {
    HostCtx<M, PP> ctx(K::NT);
    memset(smem.data(), 0xA5, smem.size());
    chain_tile<M, R, MT, PP>(ctx, p, smem.data(), (int)wi);
  }
  return (int)work.size();
}

extern "C" int emu_chain(int M, int m, int r, int dct, int C, int n_rec, const long long* Ts, const float* pcm,
                         const long long* pcm_off, float* out, const long long* out_off, const double* h,
                         const double* g, const double* w_re_im, int gain, int chunk, int fast) {
  // fast & 1: the compile-time-m instantiation (MT = m) where one exists, like the library's dispatch;
  // fast & 2: two frame pairs per warp (PP = 2)
#define ARGS m, dct, C, n_rec, Ts, pcm, pcm_off, out, out_off, h, g, w_re_im, gain, chunk
#define CASE(MM, RR) \
  if (M == MM && (1 << r) == RR) { \
    if (fast & 2) { \
      if ((fast & 1) && m == 2) return run_chain<MM, RR, 2, 2>(ARGS); \
      if ((fast & 1) && m == 4) return run_chain<MM, RR, 4, 2>(ARGS); \
      return run_chain<MM, RR, 0, 2>(ARGS); \
    } \
    if ((fast & 1) && m == 2) return run_chain<MM, RR, 2, 1>(ARGS); \
    if ((fast & 1) && m == 4) return run_chain<MM, RR, 4, 1>(ARGS); \
    return run_chain<MM, RR, 0, 1>(ARGS); \
  }
  BTK_EMU_CASES(CASE)
#undef CASE
  return -1;
}

template <int M, int R, int MT>
static int run_analysis(int m, int dct, int C, long long T, const float* pcm, float* snap, const double* h, int chunk) {
  typedef ChainCfg<M, R, MT> K;
  int r = 0; for (int x = R; x > 1; x >>= 1) r++;
  BankGeom geo(M, m, r, dct);
  const int Cpad = (C + K::CG - 1) / K::CG * K::CG;
  std::vector<cf> twa, twb; std::vector<float> hf;
  build_fft_tables(M, twa, twb);
  build_analysis_taps(h, M, m, R, hf);
  std::vector<RecDesc> recs(1);
  recs[0].pcm_off = 0; recs[0].out_off = 0; recs[0].T = (int)T; recs[0].nblk = geo.analysis_frames(T);
  std::vector<WorkItem> work;
  build_work(recs, chunk, work);
  AnalysisParams p;
  p.pcm = pcm; p.snap = reinterpret_cast<cf*>(snap); p.recs = recs.data(); p.work = work.data();
  p.taps_h = hf.data(); p.twa = twa.data(); p.twb = twb.data(); p.C = C; p.Cpad = Cpad; p.m = m; p.laN = geo.laN;
  p.cg_slices = Cpad / K::CG > 1 ? 2 : 1;
  const ChainSmem L = chain_smem_layout<M, R>(m);
  std::vector<unsigned char> smem(L.total + 64);
  for (size_t wi = 0; wi < work.size() * p.cg_slices; wi++) {
    HostCtx<M> ctx(K::NT);
    memset(smem.data(), 0xA5, smem.size());
    analysis_tile<M, R, MT>(ctx, p, smem.data(), (int)wi);
  }
  return recs[0].nblk;
}

template <int M, int R, int MT>
static int run_synthesis(int m, int dct, int F, const float* Y, float* out, const double* g, int gain, int chunk) {
  typedef ChainCfg<M, R, MT> K;
  int r = 0; for (int x = R; x > 1; x >>= 1) r++;
  BankGeom geo(M, m, r, dct);
  std::vector<cf> twa, twb; std::vector<float> gp;
  build_fft_tables(M, twa, twb);
  build_synthesis_taps(g, M, m, gp);
  std::vector<RecDesc> recs(1);
  recs[0].pcm_off = 0; recs[0].out_off = 0; recs[0].T = F; recs[0].nblk = geo.synthesis_frames(F);
  std::vector<WorkItem> work;
  build_work(recs, chunk, work);
  SynthesisParams p;
  p.Y = reinterpret_cast<const cf*>(Y); p.out = out; p.recs = recs.data(); p.work = work.data();
  p.taps_g = gp.data(); p.twa = twa.data(); p.twb = twb.data(); p.m = m; p.pd_s = geo.pd_s; p.gain = gain;
  const ChainSmem L = chain_smem_layout<M, R>(m);
  std::vector<unsigned char> smem(L.total + 64);
  for (size_t wi = 0; wi < work.size(); wi++) {
    HostCtx<M> ctx(K::NT);
    memset(smem.data(), 0xA5, smem.size());
    synthesis_tile<M, R, MT>(ctx, p, smem.data(), (int)wi);
  }
  return recs[0].nblk;
}

extern "C" int emu_analysis(int M, int m, int r, int dct, int C, long long T, const float* pcm, float* snap,
                            const double* h, int chunk, int fast) {
#define CASE(MM, RR) if (M == MM && (1 << r) == RR) { \
    if (fast && m == 2) return run_analysis<MM, RR, 2>(m, dct, C, T, pcm, snap, h, chunk); \
    if (fast && m == 4) return run_analysis<MM, RR, 4>(m, dct, C, T, pcm, snap, h, chunk); \
    return run_analysis<MM, RR, 0>(m, dct, C, T, pcm, snap, h, chunk); }
  BTK_EMU_CASES(CASE)
#undef CASE
  return -1;
}

extern "C" int emu_synthesis(int M, int m, int r, int dct, int F, const float* Y, float* out, const double* g,
                             int gain, int chunk, int fast) {
#define CASE(MM, RR) if (M == MM && (1 << r) == RR) { \
    if (fast && m == 2) return run_synthesis<MM, RR, 2>(m, dct, F, Y, out, g, gain, chunk); \
    if (fast && m == 4) return run_synthesis<MM, RR, 4>(m, dct, F, Y, out, g, gain, chunk); \
    return run_synthesis<MM, RR, 0>(m, dct, F, Y, out, g, gain, chunk); }
  BTK_EMU_CASES(CASE)
#undef CASE
  return -1;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised chain (csrc/chain_ws.cuh): the compute side runs as above; acquire() runs the producer's fill routine
// for that stage inline (every producer thread in turn), so the staged windows, the stage rotation, the aliasing of the
// v frames with the last stage and the shared exchange buffers are all exercised.  Cluster mode needs concurrently
// running CTAs and is covered by the GPU tests only.
template <int M, int PP> struct HostCtxWS {
  std::vector<ChainThreadState<M, PP> > ts;
  int nt;
  // overlap-add warps (chain_ws.cuh): the host runs their program inline where the transform warps publish an iteration;
  // "tensor memory" is an array [slot][transform warp][lane][value]
  static constexpr int NVAL = 2 * FFTGeom<M>::V;
  bool syn = false;
  int nst = 0, itc = 0;
  std::vector<float>* tm = nullptr;
  WsSynState* sst = nullptr;
  explicit HostCtxWS(int n) : ts(n), nt(n) {}
  bool syn_ok(const ChainParams&) const { return syn; }
  HostCtxWS sub() const { HostCtxWS s(nst); s.tm = tm; return s; }
  template <class F> void syn_begin_segment(F f) { HostCtxWS s = sub(); f(s, *sst); }
  int v_slot() const { return itc & 1; }
  void v_acquire(int) {}
  void tmem_store(int slot, int tid, const float* vals) {
    float* d = tm->data() + ((size_t)(slot * 8 + (tid >> 5)) * 32 + (tid & 31)) * NVAL;
    for (int i = 0; i < NVAL; i++) d[i] = vals[i];
  }
  template <class F> void v_publish(int slot, F f) { (void)slot; HostCtxWS s = sub(); f(s, *sst); itc++; }
  template <int N> void tmem_load(int slot, int w, int col0, int tid, float* vals) const {
    const float* d = tm->data() + ((size_t)(slot * 8 + w) * 32 + (tid & 31)) * NVAL + col0;
    for (int i = 0; i < N; i++) vals[i] = d[i];
  }
  void v_release(int) {}
  // second channel of a lane's pair, parked between the rounds of a stage: [thread][column]
  std::vector<float>* zb = nullptr;
  int zbw = 0;
  template <int NPKV> void zb_park_n(int tid, int step, const float* pk) { for (int i = 0; i < NPKV; i++) (*zb)[(size_t)tid * zbw + step * NPKV + i] = pk[i]; }
  void zb_parked() {}
  template <int PC> void zb_fetch_n(int tid, int col0, float* vals) { for (int i = 0; i < PC; i++) vals[i] = (*zb)[(size_t)tid * zbw + col0 + i]; }
  template <class F> void par(F f) { for (int t = 0; t < nt; t++) f(t, ts[t]); }
  void sync() {}
  void syncwarp() {}
  template <class F> void acquire(int, int, F fill) { fill(); }
  void release(int) {}
  void wait_tables() {}
  int cl_rank() const { return 0; }
  void cl_ready() {}
  void cl_expect(int, unsigned) {}
  void cl_send16(void*, int, float4, int) {}
  void cl_wait(int) {}
};

template <int M, int R, int MT, int PP>
static int run_chain_ws(int m, int dct, int C, int n_rec, const long long* Ts, const float* pcm, const long long* pcm_off,
                        float* out, const long long* out_off, const double* h, const double* g, const double* w_re_im,
                        int gain, int chunk, int syn) {
  typedef WsCfg<M, R, MT, PP> K;
  typedef FFTTables<M> FT;
  int r = 0; for (int x = R; x > 1; x >>= 1) r++;
  BankGeom geo(M, m, r, dct);
  const int Cpad = (C + K::CG - 1) / K::CG * K::CG;
  std::vector<cf> twa, twb, gam;
  std::vector<float> gp, hf;
  build_fft_tables(M, twa, twb);
  build_synthesis_taps(g, M, m, gp);
  build_analysis_taps(h, M, m, R, hf);
  std::vector<zd> w((size_t)geo.B * C);
  for (size_t i = 0; i < w.size(); i++) w[i] = zd(w_re_im[2 * i], w_re_im[2 * i + 1]);
  build_chain_weight_table(w.data(), M, C, Cpad, gam);
  std::vector<RecDesc> recs(n_rec);
  for (int i = 0; i < n_rec; i++) {
    recs[i].pcm_off = pcm_off[i]; recs[i].out_off = out_off[i]; recs[i].T = (int)Ts[i]; recs[i].nblk = geo.chain_frames(Ts[i]);
  }
  std::vector<WorkItem> work;
  build_work(recs, chunk, work);
  ChainParams p;
  p.pcm = pcm; p.out = out; p.recs = recs.data(); p.work = work.data();
  p.taps_h = hf.data(); p.taps_g = gp.data(); p.wts = gam.data(); p.wts_stride = 0; p.one_cta = 0; p.no_prefetch = 0; p.twa = twa.data(); p.twb = twb.data();
  p.C = C; p.Cpad = Cpad; p.m = m; p.pd_s = geo.pd_s; p.laN = geo.laN; p.gain = gain; p.cluster = 1;
  p.tmaps = nullptr; p.tma_rows = 0; p.item_begin = nullptr; p.item_q = 0; p.item0 = 0; p.n_items = 0; p.n_rec = n_rec;
  p.cta_begin = nullptr; p.cta_n = 0;
  p.no_syn = syn ? 0 : 1;
  const WsSmem S = ws_smem_layout<M, R, PP>(m);
  if (S.total > 227 * 1024) return -2;
  std::vector<unsigned char> smem(S.total + 64, 0xA5);
  for (size_t wi = 0; wi < work.size(); wi++) {
    HostCtxWS<M, PP> ctx(K::NT);
    std::vector<float> tm((size_t)2 * 8 * 32 * HostCtxWS<M, PP>::NVAL, 1e30f);
    WsSynState sst;
    sst.rs = 0;
    std::vector<float> zb((size_t)K::NT * K::ZBW, 1e30f);
    ctx.syn = syn != 0; ctx.nst = K::NST; ctx.tm = &tm; ctx.sst = &sst; ctx.zb = &zb; ctx.zbw = K::ZBW;
    memset(smem.data(), 0xA5, smem.size());   // poison: every read must have been written
    // what the bulk copies of the producer bring at CTA start
    memcpy(smem.data() + S.L.taps, hf.data(), (size_t)K::D * S.L.TS * 4);
    memcpy(smem.data() + S.L.twa, twa.data(), (size_t)FT::TWA_WORDS * 8);
    if (FT::TWB_WORDS) memcpy(smem.data() + S.L.twb, twb.data(), (size_t)FT::TWB_WORDS * 8);
    int g = 0;
    chain_ws_compute<M, R, MT, PP>(ctx, p, smem.data(), work[wi], recs[work[wi].rec], g);
  }
  return (int)work.size();
}

extern "C" int emu_chain_ws(int M, int m, int r, int dct, int C, int n_rec, const long long* Ts, const float* pcm,
                            const long long* pcm_off, float* out, const long long* out_off, const double* h,
                            const double* g, const double* w_re_im, int gain, int chunk, int fast) {
#define ARGS m, dct, C, n_rec, Ts, pcm, pcm_off, out, out_off, h, g, w_re_im, gain, chunk, (fast & 4) ? 0 : 1
#define CASE(MM, RR) \
  if (M == MM && (1 << r) == RR) { \
    if ((fast & 2) && MM <= 256) { \
      constexpr int PP = MM <= 256 ? 2 : 1; \
      if ((fast & 1) && m == 2) return run_chain_ws<MM, RR, 2, PP>(ARGS); \
      if ((fast & 1) && m == 4) return run_chain_ws<MM, RR, 4, PP>(ARGS); \
      return run_chain_ws<MM, RR, 0, PP>(ARGS); \
    } \
    if ((fast & 1) && m == 2) return run_chain_ws<MM, RR, 2, 1>(ARGS); \
    if ((fast & 1) && m == 4) return run_chain_ws<MM, RR, 4, 1>(ARGS); \
    return run_chain_ws<MM, RR, 0, 1>(ARGS); \
  }
  BTK_EMU_CASES(CASE)
#undef CASE
#undef ARGS
  return -1;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent schedule of the warp-specialised chain (chain_ws.cuh::WsSegs): the segments CTA `cta` of `ncta` walks for a
// launch of the recordings [r0, r1) of a batch with nblk[i] output frames each.  Writes (rec, j0, nj) triples.
extern "C" int emu_ws_segments(int n_rec, const int* nblk, int W, int r0, int r1, int cta, int ncta, int* out, int cap) {
  std::vector<RecDesc> recs(n_rec);
  std::vector<int> prefix(n_rec + 1, 0);
  for (int i = 0; i < n_rec; i++) {
    recs[i].pcm_off = 0; recs[i].out_off = 0; recs[i].T = 0; recs[i].nblk = nblk[i];
    prefix[i + 1] = prefix[i] + (nblk[i] + W - 1) / W;
  }
  ChainParams p = ChainParams();
  p.recs = recs.data(); p.item_begin = prefix.data(); p.item_q = W; p.item0 = prefix[r0]; p.n_items = prefix[r1] - prefix[r0];
  p.n_rec = n_rec;
  WsSegs segs(p, cta, ncta);
  WorkItem wk;
  int n = 0;
  while (segs.next(wk)) {
    if (n < cap) { out[3 * n] = wk.rec; out[3 * n + 1] = wk.j0; out[3 * n + 2] = wk.nj; }
    n++;
  }
  return n;
}

// Iteration-balanced CTA boundaries of the persistent schedule (host_tables.h::balance_ctas) and the segments CTA `cta`
// walks under them.  begin gets ncta + 1 item indices; returns the iteration budget B.
extern "C" int emu_ws_balance(int n_rec, const int* nblk, int q, int W, int H, int ncta, int* begin_out) {
  std::vector<RecDesc> recs(n_rec);
  std::vector<int> prefix(n_rec + 1, 0);
  for (int i = 0; i < n_rec; i++) {
    recs[i].pcm_off = 0; recs[i].out_off = 0; recs[i].T = 0; recs[i].nblk = nblk[i];
    prefix[i + 1] = prefix[i] + (nblk[i] + q - 1) / q;
  }
  std::vector<int> begin;
  const int B = balance_ctas(recs, prefix, q, W, H, ncta, begin);
  for (int i = 0; i <= ncta; i++) begin_out[i] = begin[i];
  return B;
}
extern "C" int emu_ws_segments_balanced(int n_rec, const int* nblk, int q, const int* begin, int cta, int ncta, int* out, int cap) {
  std::vector<RecDesc> recs(n_rec);
  std::vector<int> prefix(n_rec + 1, 0);
  for (int i = 0; i < n_rec; i++) {
    recs[i].pcm_off = 0; recs[i].out_off = 0; recs[i].T = 0; recs[i].nblk = nblk[i];
    prefix[i + 1] = prefix[i] + (nblk[i] + q - 1) / q;
  }
  ChainParams p = ChainParams();
  p.recs = recs.data(); p.item_begin = prefix.data(); p.item_q = q; p.item0 = 0; p.n_items = prefix[n_rec]; p.n_rec = n_rec;
  p.cta_begin = begin; p.cta_n = ncta;
  WsSegs segs(p, cta, ncta);
  WorkItem wk;
  int n = 0;
  while (segs.next(wk)) {
    if (n < cap) { out[3 * n] = wk.rec; out[3 * n + 1] = wk.j0; out[3 * n + 2] = wk.nj; }
    n++;
  }
  return n;
}
