// chain_ws.cuh -- the fused hot path (chain_tile.cuh) as a warp-specialised producer / consumer pipeline.
//
// Same arithmetic, same framing, same register programs as chain_tile.cuh (reference: modulated/modulated.cc:412-516,
// 595-664; beamformer/beamformer.cc:1137-1200, 2583-2635).  What changes is WHO moves the samples:
//
//   * NW compute warps never touch global PCM.  They wait on an mbarrier for a staged window of CG = 4 channels, run
//     polyphase -> backward transform -> weight accumulate on it, and hand the stage back.  Between channel groups there
//     is no CTA barrier any more: warps drift against each other by up to one stage.
//   * one producer warpgroup (4 warps, its registers given to the compute warps with setmaxnreg) walks the same
//     sequence of (iteration, channel group) windows one stage AHEAD: coalesced 16-byte loads of interleaved PCM,
//     register transpose, 16-byte shared stores into the residue-major window (stage_window's scheme), and ONE
//     cp.async.bulk per stage for the 4 channels' weight rows, completing on the stage's "full" mbarrier.
//   * NS = 2 stages.  The window of the next channel group (or of the next iteration's first group) lands while the
//     transforms of the current one run, so global-load latency never sits on a transform warp.
//   * the window arrives by tensor copies (cp.async.bulk.tensor, one thread) in the layout of the input where the
//     batch allows it (WsCfg::RAW; kern_ws.cuh); the producer warps with register loads are the fallback.
//   * synthesis side, tensor-copy mode: the two v frames of every frame pair leave the transform warps through TENSOR
//     MEMORY and an overlap-add warpgroup turns them into output frames (ws_syn_mode, chain_ws_synth_iter below) -- the
//     transform warps run from one iteration into the next without a CTA barrier.  Otherwise the transform warps keep
//     the synthesis side: the current-v frames alias the window of the LAST stage of an iteration where they fit, and
//     that stage is handed back after the overlap-add instead of after the last transform.
//   * persistent schedule: a CTA walks a contiguous share of the launch as per-recording segments (WsSegs below).
//
// Thread-block clusters (channel split): with p.cluster = S > 1 the S CTAs of a cluster share one work item and CTA
// `rank` takes channel groups [rank * ncg, (rank + 1) * ncg).  A row of interleaved PCM is then read by S CTAs, each
// touching only its own 16-byte granules, instead of being walked C/4 times by one CTA out of L2 -- the L2 footprint of
// all windows in flight drops by S and every PCM byte is fetched from HBM once.  The partial
// beamformer outputs G of the ranks are summed through distributed shared memory: every rank writes the partial of
// frame pair P into the receive buffer of rank P mod S, that rank adds them up, runs the forward transform and
// broadcasts the two real v frames to all ranks; every rank keeps the complete v history and emits its D/S slice of
// each output frame.  The transfers are 16-byte st.async stores that complete bytes on the receiver's mbarrier; one
// mbarrier rendezvous of the compute warps per iteration tells the peers that a CTA may be written into.  The
// producer warps take no part in it and keep running ahead.
//
// The tile program is written against a context like chain_tile.cuh, so the CPU-only tests run the very same code
// sequentially (tests/emu): ctx.acquire(stage, parity, fill) waits for the producer on the device and runs `fill`
// (the producer's per-thread routine for that stage, for every producer thread) inline on the host.
#pragma once

#include "chain_tile.cuh"

#ifndef BTK_WS_ILV
#define BTK_WS_ILV 1      // A/B: 0 keeps the lane groups in runs of 16 lanes
#endif
#ifndef BTK_WS_DUAL
#define BTK_WS_DUAL 1     // A/B: 0 windows one channel per pass also where the overlap-add warps (tensor memory) are there
#endif

namespace btk {

template <int M_, int R_, int MT_ = 0, int PP_ = 1>
struct WsCfg {
  typedef FFTGeom<M_> G;
  static constexpr int M = M_, R = R_, D = M_ / R_;
  static constexpr int MT = MT_;
  static constexpr int PP = PP_;
  // eight compute warps (two per scheduler) wherever the exchange buffers allow; the 32-lane transforms of M = 1024
  // keep four
  static constexpr int NW = M_ >= 1024 ? 4 : 8;
  static constexpr int NT = NW * 32;
  static constexpr int NPW = 4;                // producer warps: one warpgroup (setmaxnreg works on warpgroups)
  static constexpr int NPT = NPW * 32;
  static constexpr int FW = 2 * PP_;
  static constexpr int LV = (FW % 4 == 0) ? 4 : 2;
  static constexpr int W = FW * NW;
  static constexpr int CG = 4;
  // RAW: the staged window keeps the layout of the interleaved input, [time step][CG] (16 bytes per time step) -- the
  // producer's 16-byte load goes out as ONE 16-byte store, no register transpose; the polyphase reads single floats,
  // D * CG floats apart, instead of vectors of consecutive blocks.  (This is also the layout a tensor bulk copy delivers.)
  static constexpr bool RAW = BTK_WS_RAW != 0;
  static constexpr int NG = G::NG;
  // exchange buffers per lane group: the two frame pairs of M = 256 take turns on one (analysis_round), which is what
  // lets two 4-channel stages of the 32-frame window fit next to the buffers of eight warps
  static constexpr int XS = (PP_ == 2 && G::Rb == 1 && G::Ra <= 16 && G::NG == 2) ? 1 : PP_;
  static constexpr int E = G::Ra / R_;
  // Lane <-> (lane group, lane inside the group).  With two groups of 16 lanes the groups are INTERLEAVED in runs of eight
  // lanes (lanes 0-7: group 0, 8-15: group 1, 16-23: group 0, 24-31: group 1), so that a half-warp -- the unit an 8-byte
  // shared-memory access is served in -- holds eight lanes of EACH group: the two groups' exchange buffers are then kept 16
  // banks apart (XPAD), and an access that takes two channels per lane out of the raw window uses whole 16-byte granules.
  static constexpr bool ILV = BTK_WS_ILV != 0 && G::NG == 2 && G::L == 16;
  static BTK_HD int lane_grp(int lane) { return ILV ? (lane >> 3) & 1 : lane / G::L; }
  static BTK_HD int lane_gl(int lane) { return ILV ? (lane & 7) | ((lane >> 4) << 3) : lane % G::L; }
  static constexpr int XPAD = ILV ? 8 : 0;
  static constexpr int FPT_RAW = (W * D) / NT;
  static constexpr int FPT = FPT_RAW >= 8 ? 8 : (FPT_RAW >= 4 ? 4 : (FPT_RAW >= 2 ? 2 : 1));
  static constexpr int NS = 2;                 // stages
  // Overlap-add warps (tensor-copy mode): one warpgroup next to the producer's takes the synthesis side's polyphase with
  // g, the overlap-add and the output stores off the transform warps.  The v frames of an iteration travel through TENSOR
  // MEMORY: a transform warp parks the 2 V values per lane its forward transform leaves in registers (tcgen05.st, its own
  // 32 lanes of the 128, NVAL columns), the overlap-add warp of the same lane quarter (warp index mod 4) takes them out
  // chunk by chunk (tcgen05.ld) into a small ring of frames in shared memory.  256 KB of tensor memory sit idle in
  // this kernel otherwise; shared memory has no 32 KB left for a second set of v frames.
  static constexpr int NSY = 4;
  static constexpr int NST = NSY * 32;
  static constexpr int NVAL = 2 * G::V;                                  // floats a lane parks per iteration
  static constexpr int CHF = 4;                                          // frames per chunk of the overlap-add warps
  // Tensor memory also holds, for the time between the two rounds of a stage, the windowed frames of the SECOND channel
  // of each lane's channel pair (chain_tile.cuh::polyphase_pairs2): ZBW = 2 PP V floats per lane and transform warp
  static constexpr int ZBW = 2 * PP_ * G::V;
  static constexpr int NPK = 2 * PP_ * R_;                               // floats parked per residue step
  static constexpr int ZB0 = 4 * NVAL;                                   // first column of that area (2 warps per lane quarter)
  static constexpr int TM_NEED = 4 * NVAL + 2 * ZBW;                     // v frames: 2 slots x 2 warps per lane quarter x NVAL
  static constexpr int TM_COLS = TM_NEED <= 32 ? 32 : (TM_NEED <= 64 ? 64 : (TM_NEED <= 128 ? 128 : (TM_NEED <= 256 ? 256 : 512)));
  static constexpr bool DUAL_OK = NG == 2 && CG == 4 && RAW && MT_ > 0 && TM_NEED <= 512 && (NPK == 8 || NPK == 16 || NPK == 32);
  static_assert(G::Ra % R_ == 0, "decimation factor must divide the first radix");
  static_assert(CG % NG == 0, "channel group must be a multiple of the lane groups per warp");
  static_assert(W % FPT == 0, "frames per thread must divide the iteration");
};

// what the overlap-add warps see of the configuration: their own thread count
template <int M_, int R_, int MT_ = 0, int PP_ = 1>
struct WsSynCfg : WsCfg<M_, R_, MT_, PP_> {
  typedef WsCfg<M_, R_, MT_, PP_> B;
  static constexpr int NT = B::NST;
};

struct WsSmem {
  ChainSmem L;          // TS, TV, NB, SB, CS, H + taps / twa / twb offsets (the fields the shared tile pieces read)
  int bars;             // mbarriers: full[NS], empty[NS], tables, cluster barriers
  int stage0;           // first stage; a stage = window of CG channels, then their weight rows
  int stage_bytes;
  int wts_off;          // offset of the weight rows inside a stage
  int xbuf, vhist, vcur;
  int rbuf;             // cluster receive buffer (aliases the exchange buffers: they are idle during the reduction)
  int valias;           // current-v frames alias the window of a stage
  int ring;             // overlap-add warps: ring of NRF v frames (aliases the history frames: they own both), FS floats apart
  int NRF, FS;          // 0 = no room for the ring: the transform warps keep the synthesis side
  int total;
};

enum { WS_BAR_FULL = 0, WS_BAR_EMPTY = 2, WS_BAR_TABLES = 4, WS_BAR_READY = 5, WS_BAR_RX0 = 6, WS_BAR_RX1 = 7, WS_BAR_VFULL = 8,
       WS_BAR_VEMPTY = 10, WS_NBARS = 12 };

template <int M_, int R_, int PP_>
BTK_HD constexpr WsSmem ws_smem_layout(int m) {
  typedef WsCfg<M_, R_, 0, PP_> K;
  typedef FFTTables<M_> FT;
  WsSmem s = WsSmem();
  const int mR = m * R_;
  s.L.TV = tap_vec(mR);
  s.L.TS = tap_stride(mR);
  s.L.NB = K::W - 1 + mR;
  int sb = 0;
  if (K::LV == 4) {
    sb = (s.L.NB + 3) & ~3;
    if (((sb / 4) & 1) != 0) sb += 4;
  } else {
    sb = (s.L.NB + 1) & ~1;
    if (((sb / 2) & 1) == 0) sb += 2;
  }
  s.L.SB = sb;
  int cs = K::D * sb;
  if (K::G::L < 16) cs += (16 - (cs & 31) + 32) & 31;
  s.L.CS = cs;
  s.L.H = mR - 1;
  int off = 0;
  s.bars = off;   off += WS_NBARS * 8;                  off = (off + 127) & ~127;
  s.L.taps = off; off += K::D * s.L.TS * 4;             off = (off + 15) & ~15;
  s.L.twa = off;  off += FT::TWA_WORDS * 8;             off = (off + 15) & ~15;
  s.L.twb = off;  off += FT::TWB_WORDS * 8;             off = (off + 127) & ~127;
  // the raw window is exactly NB D time steps of CG channels; the residue-major one pads its rows
  const int win = K::RAW ? s.L.NB * K::D * K::CG * 4 : K::CG * cs * 4;
  s.wts_off = (win + 127) & ~127;
  s.stage_bytes = (s.wts_off + K::CG * M_ * 8 + 127) & ~127;
  s.stage0 = off; off += K::NS * s.stage_bytes;
  s.xbuf = off;   off += K::NW * K::NG * (K::XS * K::G::XBUF + K::XPAD) * 8;   off = (off + 15) & ~15;
  s.vhist = off;  off += (s.L.H > 0 ? s.L.H : 1) * M_ * 4;          off = (off + 15) & ~15;
  s.valias = (K::W * M_ * 4 <= win) ? 1 : 0;
  s.vcur = off;
  if (!s.valias) off += K::W * M_ * 4;
  {
    // ring of the overlap-add warps: the H history frames + one chunk, in whole chunks; frames 8 floats further apart
    // than M so that the two lane groups of a warp (frames two apart) store into different banks
    const int nrf = ((s.L.H + K::CHF - 1) / K::CHF + 1) * K::CHF, fs = M_ + 8;
    const int ring_end = ((s.vhist + nrf * fs * 4) + 15) & ~15;
    s.ring = s.vhist; s.NRF = 0; s.FS = fs;
    if (K::NG == 2 && K::W % K::CHF == 0 && nrf / K::CHF <= 8 && (ring_end > off ? ring_end : off) <= 227 * 1024 - 1024) {
      s.NRF = nrf;
      if (ring_end > off) off = ring_end;
    }
  }
  s.rbuf = s.xbuf;
  s.L.xs = s.stage0; s.L.wts = s.stage0 + s.wts_off; s.L.xbuf = s.xbuf; s.L.vhist = s.vhist;
  s.total = off;
  s.L.total = off;
  return s;
}

// cluster receive buffer: NW * PP pairs in total, [src rank][owned pair] blocks of V * L complex words
template <class K>
BTK_HD constexpr bool ws_cluster_ok(int S) {
  return S >= 1 && (K::NW * K::PP) % S == 0 && K::D % S == 0 &&
         K::NW * K::PP * K::M * 8 <= K::NW * K::NG * K::XS * K::G::XBUF * 8 && K::NG <= 2;
}

// Persistent schedule.  The launch is a run of items (an item = W consecutive output frames of one recording, numbered
// across the recordings of the batch by the prefix sums p.item_begin); CTA `cta` of `ncta` takes the contiguous share
// [n cta / ncta, n (cta + 1) / ncta) and walks it as SEGMENTS: the items of one recording that follow each other are one
// WorkItem, so the synthesis history is warmed up once per segment and not once per item, every CTA runs the same number
// of iterations to within two, the tables are loaded once per SM and the producer runs ahead across segment boundaries.
// An item is a few frames (W / 8), not an iteration: a segment of nj frames costs ceil((nj + H) / W) iterations (H = m R - 1
// warm-up frames), so equal ITEM counts leave the CTAs one or two iterations apart -- 27 against 25.3 at cfg2.  For a
// launch of the whole batch the host therefore hands over the first item of every CTA (p.cta_begin,
// host_tables.h::balance_ctas: the smallest iteration budget under which a greedy walk needs no more CTAs than there are).
// With p.item_begin == NULL the CTA has exactly one segment, p.work[cta] (the chunk list of the first sessions).
struct WsSegs {
  const ChainParams& p;
  int cta, i, i1, r;
  bool one_shot;
  BTK_HD WsSegs(const ChainParams& p_, int cta_, int ncta) : p(p_), cta(cta_), i(0), i1(0), r(0), one_shot(false) {
    if (!p.item_begin) { one_shot = true; return; }
    const long long n = p.n_items;
    if (p.cta_begin) {
      i = p.cta_begin[cta];
      i1 = p.cta_begin[cta + 1];
    } else {
      i = p.item0 + (int)(n * cta / ncta);
      i1 = p.item0 + (int)(n * (cta + 1) / ncta);
    }
    if (i >= i1) return;
    // last recording whose first item is <= i: item_begin is non-decreasing, item_begin[0] = 0
    int lo = 0, hi = p.n_rec;                                    // item_begin[n_rec] (the total) > i
    while (hi - lo > 1) {
      const int mid = (lo + hi) / 2;
      if (p.item_begin[mid] <= i) lo = mid; else hi = mid;
    }
    r = lo;
  }
  BTK_HD bool next(WorkItem& wk) {
    if (one_shot) {
      if (i1) return false;
      i1 = 1;
      wk = p.work[cta];
      return true;
    }
    if (i >= i1) return false;
    while (p.item_begin[r + 1] <= i) r++;                        // recordings without frames have no items
    const int b = p.item_begin[r], e = p.item_begin[r + 1] < i1 ? p.item_begin[r + 1] : i1;
    const long long j1 = (long long)(e - b) * p.item_q;
    const int nblk = p.recs[r].nblk;
    wk.rec = r;
    wk.j0 = (i - b) * p.item_q;
    wk.nj = (int)(j1 < nblk ? j1 : nblk) - wk.j0;
    i = e;
    return true;
  }
};

// The window one producer pass stages: iteration `it`, channel group `cgi` of this CTA.
struct WsWalk {
  int a_start, n_it, ncg, cg_base;
  BTK_HD int groups() const { return n_it * ncg; }
};

// One producer thread's share of a stage: CG channels [cg0, cg0 + CG) of the window starting at sample t_lo,
// pcm [t][C] -> s_xs[c][t mod D][t div D] (residue-major, see chain_tile.cuh::stage_window for the layout and the
// register transpose).  A batch is TB tasks (= TB * LV 16-byte loads in flight per thread); the loads and the stores of
// a batch are separate calls so that the device producer can issue the first batch of a stage BEFORE it waits for the
// stage to be handed back (the loads only need registers), and so that the load latency of that batch hides behind the
// wait.
// (A lean variant of the load loop without the per-sample range tests for interior windows measured SLOWER on every
// shape -- cfg2 0.3614 -> 0.3657 ms, cfg3 0.8235 -> 0.8747 ms, cfg4 1.887 -> 2.003 ms: its loads leave back to back, and
// a burst of 32-sector requests in the LSU queue delays the shared-memory accesses of the transform warps.)
template <class K>
struct WsFill {
  const float* pcm_cg;
  float* s_xs;
  int t0, T, C, cg0, ntask, nbatch, SB, CS, NB;
  bool v4;
  BTK_HD WsFill(const ChainSmem& L, float* xs, const float* pcm, int C_, int T_, long long t_lo, int cg0_, bool vec4, int tb)
      : pcm_cg(pcm + cg0_), s_xs(xs), t0((int)t_lo), T(T_), C(C_), cg0(cg0_), SB(L.SB), CS(L.CS), NB(L.NB) {
    v4 = vec4 && cg0_ + K::CG <= C_;
    ntask = K::D * ((L.NB + K::LV - 1) / K::LV);
    nbatch = (ntask + K::NPT * tb - 1) / (K::NPT * tb);
  }
  template <int TB>
  BTK_HD void load(int ptid, int b, float (&x)[TB][K::LV][K::CG]) const {
    constexpr int D = K::D, LV = K::LV;
    const int task0 = ptid + b * (K::NPT * TB);
#ifdef BTK_WS_FASTFILL       // A/B knob: interior windows without the per-sample range tests (see the note above)
    if (v4 && t0 >= 0 && (long long)t0 + (long long)NB * D <= (long long)T) {
      const size_t row = (size_t)D * (unsigned)C;
      BTK_UNROLL
      for (int k = 0; k < TB; k++) {
        const int task = task0 + k * K::NPT;
        const int res = task % D, bg = task / D;
        const float* src = pcm_cg + (size_t)((unsigned)(t0 + res)) * (unsigned)C + (size_t)(bg * LV) * row;
        BTK_UNROLL
        for (int i = 0; i < LV; i++) {
          float4 q; q.x = 0.f; q.y = 0.f; q.z = 0.f; q.w = 0.f;
          if (task < ntask && bg * LV + i < NB) q = *reinterpret_cast<const float4*>(src + i * row);
          x[k][i][0] = q.x; x[k][i][1] = q.y; x[k][i][2] = q.z; x[k][i][3] = q.w;
        }
      }
      return;
    }
#endif
    BTK_UNROLL
    for (int k = 0; k < TB; k++) {
      const int task = task0 + k * K::NPT;
      const int res = task % D, bg = task / D;
      BTK_UNROLL
      for (int i = 0; i < LV; i++) {
        const int blk = bg * LV + i;
        const int t = t0 + blk * D + res;
        BTK_UNROLL
        for (int c = 0; c < K::CG; c++) x[k][i][c] = 0.f;
        if (task < ntask && blk < NB && (unsigned)t < (unsigned)T) {
          const float* src = pcm_cg + (size_t)((unsigned)t) * (unsigned)C;
          if (v4) {
#if defined(__CUDA_ARCH__) && defined(BTK_WS_LDG_NOALLOC)
            float4 q;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(src));
#else
            const float4 q = *reinterpret_cast<const float4*>(src);
#endif
            x[k][i][0] = q.x; x[k][i][1] = q.y; x[k][i][2] = q.z; x[k][i][3] = q.w;
          } else {
            BTK_UNROLL
            for (int c = 0; c < K::CG; c++) if (cg0 + c < C) x[k][i][c] = src[c];
          }
        }
      }
    }
  }
  template <int TB>
  BTK_HD void store(int ptid, int b, const float (&x)[TB][K::LV][K::CG]) const {
    constexpr int D = K::D, LV = K::LV;
    const int task0 = ptid + b * (K::NPT * TB);
    BTK_UNROLL
    for (int k = 0; k < TB; k++) {
      const int task = task0 + k * K::NPT;
      if (task < ntask) {
        const int res = task % D, bg = task / D;
        if (K::RAW) {
          BTK_UNROLL
          for (int i = 0; i < LV; i++) {
            const int blk = bg * LV + i;
            if (blk < NB) {
              float4 v; v.x = x[k][i][0]; v.y = x[k][i][1]; v.z = x[k][i][2]; v.w = x[k][i][3];
              *reinterpret_cast<float4*>(s_xs + (size_t)(blk * D + res) * K::CG) = v;
            }
          }
          continue;
        }
        float* dst = s_xs + xs_off<LV>(res, bg, SB);
        BTK_UNROLL
        for (int c = 0; c < K::CG; c++) {
          if (LV == 4) {
            float4 v; v.x = x[k][0][c]; v.y = x[k][1][c]; v.z = x[k][LV > 2 ? 2 : 0][c]; v.w = x[k][LV - 1][c];
            *reinterpret_cast<float4*>(dst + c * CS) = v;
          } else {
            float2 v; v.x = x[k][0][c]; v.y = x[k][1][c];
            *reinterpret_cast<float2*>(dst + c * CS) = v;
          }
        }
      }
    }
  }
};

// the whole share of one producer thread in one call (host emulation; the device producer drives WsFill itself)
template <class K, int TB>
BTK_HD void ws_fill_thread(int ptid, const ChainSmem& L, float* s_xs, const float* pcm, int C, int T, long long t_lo,
                           int cg0, bool vec4) {
  const WsFill<K> f(L, s_xs, pcm, C, T, t_lo, cg0, vec4, TB);
  for (int b = 0; b < f.nbatch; b++) {
    float x[TB][K::LV][K::CG];
    f.template load<TB>(ptid, b, x);
    f.template store<TB>(ptid, b, x);
  }
}

template <class K>
BTK_HD WsWalk ws_walk(const ChainParams& p, const WorkItem& wk, int H, int csz, int rank) {
  WsWalk w;
  w.a_start = wk.j0 + p.pd_s - H;
  w.n_it = (wk.nj + H + K::W - 1) / K::W;
  w.ncg = (p.Cpad / K::CG) / csz;
  w.cg_base = rank * w.ncg;
  return w;
}

// oldest sample of the window of iteration `it`: frame i = tau_base + laN needs x[(i+1) D - N .. (i+1) D - 1]
template <class K>
BTK_HD long long ws_window_start(const WsWalk& w, int it, int laN, int N) {
  const int tau_base = w.a_start + it * K::W;
  return (long long)(tau_base + laN + 1) * K::D - N;
}

// ---------------------------------------------------------------------------------------------
// Overlap-add warps.  Active when the window comes by tensor copies (the producer is then one thread of another
// warpgroup), without a cluster, with a compile-time prototype length and two lane groups per warp (M = 128 .. 512).
// Protocol per iteration (slot = iteration counter of the CTA & 1, the counter runs on across segments):
//   transform warp w   wait VEMPTY[slot] -> park NVAL floats per lane at columns (slot * 2 + w / 4) * NVAL of its lane
//                      quarter -> arrive VFULL[slot] (count NW);  nothing else: no CTA barrier, no shared memory
//   overlap-add warps  wait VFULL[slot]; for every chunk of CHF = 4 frames, in frame order: the warp(s) of the lane
//                      quarter(s) holding it copy it into the next slot of the ring, barrier, ALL overlap-add threads emit
//                      the chunk's output frames (modulated.cc:646-661) from the ring (the history is the chunks before
//                      it), barrier; after its last copy of the iteration every warp arrives on VEMPTY[slot] (count NSY)
// ---------------------------------------------------------------------------------------------
template <class K>
BTK_HD bool ws_syn_mode(const WsSmem& S, bool ctx_ok, int csz) {
  return ctx_ok && csz <= 1 && S.NRF > 0 && K::MT > 0 && K::NG == 2;
}

struct WsSynState {
  int rs;        // ring slot (chunk index mod NRF / CHF) the next chunk goes to
};

template <int M_, int R_, int MT_, int PP_, class Ctx>
BTK_HD void chain_ws_synth_begin(Ctx& sctx, unsigned char* smem, const WsSmem& S, WsSynState& st) {
  typedef WsSynCfg<M_, R_, MT_, PP_> KS;
  typedef ChainThreadState<M_, PP_> TS;
  float* ring = reinterpret_cast<float*>(smem + S.ring);
  sctx.par([&](int tid, TS&) {
    for (int i = tid; i < S.NRF * S.FS; i += KS::NT) ring[i] = 0.f;       // zero history (modulated.cc:666-674)
  });
  st.rs = 0;
  sctx.sync();
}

// One iteration of the overlap-add warps: W / CHF chunks.  sctx.tmem_load(slot, w, vals) hands a thread of the overlap-add
// warp (w mod 4) the NVAL floats lane (tid mod 32) of transform warp w parked: value 2 r = v_{f0}[q_r], 2 r + 1 =
// v_{f0 + 1}[q_r], q_r = index_of(gl, r), f0 = FW w + 2 grp for two frame pairs per warp, 2 w for one (group 0 only).
template <int M_, int R_, int MT_, int PP_, class Ctx>
BTK_HD void chain_ws_synth_iter(Ctx& sctx, const ChainParams& p, unsigned char* smem, const WsSmem& S, const WorkItem wk,
                                const RecDesc rec, int it, int slot, WsSynState& st) {
  typedef WsSynCfg<M_, R_, MT_, PP_> KS;
  typedef typename KS::G G;
  typedef ChainThreadState<M_, PP_> TS;
  constexpr int CHF = KS::CHF, NCH = KS::W / CHF, D = KS::D, mm = MT_ > 0 ? MT_ : 1;
  constexpr int WPC = CHF / KS::FW;                              // transform warps per chunk (1, or 2 with one pair per warp)
  static_assert(WPC >= 1 && CHF % KS::FW == 0, "a chunk is whole transform warps");
  const ChainSmem& L = S.L;
  const int NRS = S.NRF / CHF, FS = S.FS;
  float* ring = reinterpret_cast<float*>(smem + S.ring);
  float* out = p.out + rec.out_off;
  const int tau_base = wk.j0 + p.pd_s - L.H + it * KS::W;
  const float gf = p.gain > 0 ? (float)p.gain : 1.f;
  for (int c = 0; c < NCH; c++) {
    float* cur = ring + st.rs * CHF * FS;
    // ---- the chunk's frames out of tensor memory into the ring
    sctx.par([&](int tid, TS&) {
      const int sw = tid >> 5, lane = tid & 31, grp = KS::lane_grp(lane), gl = KS::lane_gl(lane);
      BTK_UNROLL
      for (int k = 0; k < WPC; k++) {
        const int w = c * WPC + k;
        if ((w & 3) != sw) continue;
        // in pieces of at most 32 floats: the overlap-add warps run on a small register budget
        constexpr int PC = KS::NVAL < 32 ? KS::NVAL : 32;
        float* v0 = cur + (KS::FW * k + (PP_ == 2 ? 2 * grp : 0)) * FS;
        BTK_UNROLL
        for (int h = 0; h < KS::NVAL / PC; h++) {
          float vals[PC];
          sctx.template tmem_load<PC>(slot, w, h * PC, tid, vals);
          if (PP_ == 1 && grp != 0) continue;                    // one pair per warp: lane group 0 holds it
          BTK_UNROLL
          for (int r2 = 0; r2 < PC / 2; r2++) {
            const int q = G::index_of(gl, h * (PC / 2) + r2);
            v0[q] = vals[2 * r2];
            v0[FS + q] = vals[2 * r2 + 1];
          }
        }
      }
    });
    if (c + 1 == NCH) sctx.v_release(slot);                      // this warp's last copy of the iteration is done
    sctx.sync();
    // ---- output frames jb .. jb + CHF - 1 from the ring: frame e of the chunk (e < 0: the chunks before it) sits in slot
    // (rs - k) mod NRS with k = (CHF - 1 - e) / CHF chunks back
    {
      const int jb = tau_base + c * CHF - p.pd_s;
      sctx.par([&](int tid, TS&) {
        if (jb + CHF <= wk.j0 || jb >= wk.j0 + wk.nj) return;
        const float* base[8];
        BTK_UNROLL
        for (int k = 0; k < 8; k++) base[k] = ring + ((st.rs + (NRS - (k % NRS))) % NRS) * CHF * FS;
        for (int d = tid; d < D; d += KS::NT) {
          float acc[CHF];
          BTK_UNROLL
          for (int f = 0; f < CHF; f++) acc[f] = 0.f;
          BTK_UNROLL
          for (int s = 0; s < R_; s++) {
            const int back = R_ - 1 - s, q = d + s * D;
            constexpr int NV = CHF + R_ * (mm - 1);
            float g[mm], v[NV];
            BTK_UNROLL
            for (int k = 0; k < mm; k++) g[k] = p.taps_g[k * M_ + q];
            BTK_UNROLL
            for (int i = 0; i < NV; i++) {
              const int e = i - back - R_ * (mm - 1);            // chunk-relative frame, compile time
              const int kb = e >= 0 ? 0 : (CHF - 1 - e) / CHF;
              v[i] = base[kb][(e + kb * CHF) * FS + q];
            }
            BTK_UNROLL
            for (int f = 0; f < CHF; f++) {
              float w = 0.f;
              BTK_UNROLL
              for (int k = 0; k < mm; k++) w = fmaf(g[k], v[f + R_ * (mm - 1) - R_ * k], w);
              if (jb + f - back >= 0) acc[f] += w;
            }
          }
          BTK_UNROLL
          for (int f = 0; f < CHF; f++) {
            const int j = jb + f;
            if (j >= wk.j0 && j < wk.j0 + wk.nj) out[(long long)j * D + (D - 1 - d)] = acc[f] * gf;
          }
        }
      });
    }
    sctx.sync();
    st.rs = st.rs + 1 == NRS ? 0 : st.rs + 1;
  }
}

// ---------------------------------------------------------------------------------------------
// The compute side of the tile program.  Ctx provides, besides par / sync / syncwarp of chain_tile.cuh (sync = barrier
// of the COMPUTE threads only):
//   acquire(stage, parity, fill)    wait until the producer has filled `stage`; the host context runs fill() instead
//   release(stage)                  this warp (device) / the CTA (host) is done with `stage`
//   wait_tables()                   taps and twiddles have landed
//   cluster hooks (only used when p.cluster > 1):
//     cl_rank()                       rank of this CTA in its cluster
//     cl_ready()                      barrier of the compute warps of the whole cluster
//     cl_expect(k, bytes)             this CTA will receive `bytes` on receive barrier k in this iteration
//     cl_send16(ptr, rank, v, k)      store 16 bytes at the same shared-memory address in CTA `rank`, counted on ITS barrier k
//     cl_wait(k)                      everything expected on receive barrier k has landed
// ---------------------------------------------------------------------------------------------
// SYNM: the synthesis side is with the overlap-add warps (1), with the transform warps (0) -- the device kernels are built
// once for each, so that neither carries the other's code -- or decided at run time (-1, host emulation).
template <int M_, int R_, int MT_, int PP_, class Ctx, int SYNM = -1>
BTK_HD void chain_ws_compute(Ctx& ctx, const ChainParams& p, unsigned char* smem, const WorkItem wk, const RecDesc rec, int& g) {
  typedef WsCfg<M_, R_, MT_, PP_> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_, PP_> TS;
  const int m = MT_ > 0 ? MT_ : p.m;
  const int N = M_ * m;
  const WsSmem S = ws_smem_layout<M_, R_, PP_>(m);
  const ChainSmem& L = S.L;
  const int H = L.H;
  float* s_taps = reinterpret_cast<float*>(smem + L.taps);
  cf* s_twa = reinterpret_cast<cf*>(smem + L.twa);
  cf* s_twb = reinterpret_cast<cf*>(smem + L.twb);
  cf* s_xbuf = reinterpret_cast<cf*>(smem + S.xbuf);
  float* s_vhist = reinterpret_cast<float*>(smem + S.vhist);

  const float* pcm = p.pcm + rec.pcm_off;
  float* out = p.out + rec.out_off;
  const int C = p.C;
  const int csz = SYNM == 1 ? 1 : (p.cluster > 1 ? p.cluster : 1);
  const int rank = csz > 1 ? ctx.cl_rank() : 0;
  const WsWalk walk = ws_walk<K>(p, wk, H, csz, rank);
  const bool vec4 = (C % 4 == 0) && (rec.pcm_off % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0);
  const cf* wts = p.wts + (long long)wk.rec * p.wts_stride;
  // overlap-add warps take the v frames of every iteration out of tensor memory: no synthesis side here (ws_syn_mode)
  const bool syn = SYNM >= 0 ? SYNM == 1 : ws_syn_mode<K>(S, ctx.syn_ok(p), csz);

  if (!syn) {
    ctx.par([&](int tid, TS&) {
      for (int i = tid; i < H * M_; i += K::NT) s_vhist[i] = 0.f;
    });
  } else {
    ctx.syn_begin_segment([&](auto& sctx, WsSynState& st) { chain_ws_synth_begin<M_, R_, MT_, PP_>(sctx, smem, S, st); });
  }
  ctx.wait_tables();
  ctx.sync();

  // g: stages consumed so far by this CTA (it runs on across the segments of a persistent CTA, like the producer's)
  for (int it = 0; it < walk.n_it; it++) {
    const int tau_base = walk.a_start + it * K::W;
    const long long t_lo = ws_window_start<K>(walk, it, p.laN, N);
    ctx.par([&](int, TS& ts) {
      BTK_UNROLL
      for (int r = 0; r < PP_ * G::V; r++) ts.g[r] = mk(0.f, 0.f);
    });
    int s_last = 0;
    for (int cgi = 0; cgi < walk.ncg; cgi++, g++) {
      const int st = g % K::NS;
      unsigned char* stage = smem + S.stage0 + st * S.stage_bytes;
      float* s_xs = reinterpret_cast<float*>(stage);
      const float4* s_wts = reinterpret_cast<const float4*>(stage + S.wts_off);
      const int cg0 = (walk.cg_base + cgi) * K::CG;
      ctx.acquire(st, (g / K::NS) & 1, [&]() {
        for (int ptid = 0; ptid < K::NPT; ptid++) ws_fill_thread<K, 1>(ptid, L, s_xs, pcm, C, rec.T, t_lo, cg0, vec4);
        const float4* src = reinterpret_cast<const float4*>(wts + (long long)cg0 * M_);
        float4* dst = reinterpret_cast<float4*>(stage + S.wts_off);
        for (int i = 0; i < K::CG * M_ / 2; i++) dst[i] = src[i];
      });
#ifdef BTK_EXP_NOCOMPUTE    // experiment: the producer side alone (no transforms; results are garbage)
      for (int round = 0; round < 0; round++) {
#else
      for (int round = 0; round < K::CG / K::NG; round++) {
#endif
        // two-channel mode (with the overlap-add warps, i.e. with tensor memory): lane group grp windows channels 2 grp and
        // 2 grp + 1 of the stage together in round 0 -- one 8-byte load per sample pair -- transforms the first and parks
        // the second in tensor memory until round 1; otherwise channel round * NG + grp is windowed in its own round
        const bool dual = K::DUAL_OK && syn && (SYNM == 1 ? BTK_WS_DUAL != 0 : !(p.no_syn & 2));
        if (!dual) {
          analysis_round<K>(ctx, L, s_xs, s_taps, s_xbuf, s_twa, s_twb, m, round);
        } else if constexpr (K::DUAL_OK) {
          analysis_round_fill<K>(ctx, s_xbuf, s_twa, s_twb, [&](cf* z, int tid, int grp, int gl) {
            const int warp = tid >> 5;
            if (round == 0) {
              polyphase_pairs2<K>(z, gl, s_xs + 2 * grp, warp, s_taps, L,
                                  [&](int step, const float* pk) { ctx.template zb_park_n<K::NPK>(tid, step, pk); });
              ctx.zb_parked();
            } else {
              constexpr int PC = K::ZBW < 32 ? K::ZBW : 32;
              polyphase_unpark<K>(z, [&](int col0, float* vals) { ctx.template zb_fetch_n<PC>(tid, col0, vals); });
            }
          });
        }
        ctx.par([&](int tid, TS& ts) {
          const int lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
          const float4* w4 = s_wts + (dual ? 2 * grp + round : round * K::NG + grp) * (G::V / 2) * G::L + gl;
          BTK_UNROLL
          for (int r2 = 0; r2 < G::V / 2; r2++) {
            const float4 w = w4[r2 * G::L];
            BTK_UNROLL
            for (int pp = 0; pp < PP_; pp++) {
              cfma(ts.g[pp * G::V + 2 * r2], ts.z[pp * G::V + 2 * r2], mk(w.x, w.y));
              cfma(ts.g[pp * G::V + 2 * r2 + 1], ts.z[pp * G::V + 2 * r2 + 1], mk(w.z, w.w));
            }
          }
        });
        ctx.syncwarp();
      }
      s_last = st;
      // the stage whose window the v frames alias stays with the compute warps until the overlap-add is done
      if (syn || !(S.valias && cgi + 1 == walk.ncg)) ctx.release(st);
    }
    if (syn) {
      // ---- forward transforms of this warp's pairs, the two frames of each left in registers; park them in tensor memory
      // for the overlap-add warps and go on with the next iteration
      synth_gather_pairs<K>(ctx, s_xbuf);
      synth_transform_store<K>(ctx, s_xbuf, s_twa, s_twb, (float*)nullptr, tau_base, 1, 0);
      const int slot = ctx.v_slot();
      ctx.v_acquire(slot);
      ctx.par([&](int tid, TS& ts) {
        float vals[K::NVAL];
        BTK_UNROLL
        for (int r = 0; r < G::V; r++) { vals[2 * r] = ts.g[r].x; vals[2 * r + 1] = ts.g[r].y; }
        ctx.tmem_store(slot, tid, vals);
      });
      ctx.v_publish(slot, [&](auto& sctx, WsSynState& st) {
        chain_ws_synth_iter<M_, R_, MT_, PP_>(sctx, p, smem, S, wk, rec, it, slot, st);
      });
      continue;
    }
    float* s_vcur = S.valias ? reinterpret_cast<float*>(smem + S.stage0 + s_last * S.stage_bytes)
                             : reinterpret_cast<float*>(smem + S.vcur);

    synth_gather_pairs<K>(ctx, s_xbuf);
    if (csz > 1) {
      // ---- sum the partial G of the ranks: pair P = warp * PP + pp goes to rank P % S, slot [src rank][P / S] of its receive
      // buffer, stored as [r / 2][gl][2] complex words so that the lanes of a group write consecutive 16-byte packets.
      // The exchange is transaction counted: every 16-byte remote store (st.async) completes 16 bytes on the RECEIVER's
      // mbarrier, the receiver waits for the byte count it expects -- no cluster-scope fence anywhere (a
      // release.cluster arrive compiles to MEMBAR.ALL.GPU, one per peer: measured 23 us per iteration at eight ranks).
      constexpr int PW = G::V * G::L;                      // complex words per pair
      const int own = (K::NW * K::PP) / csz;               // pairs this rank owns
      cf* rbuf = reinterpret_cast<cf*>(smem + S.rbuf);
      // every warp of every rank is past its last transform: exchange buffers idle, staged windows read, the previous
      // iteration's v frames consumed -- peers may now write into this CTA
      ctx.cl_ready();
      ctx.cl_expect(0, (unsigned)((csz - 1) * own * PW * 8));
      ctx.cl_expect(1, (unsigned)((K::NW * K::PP - own) * 2 * M_ * 4));
      ctx.par([&](int tid, TS& ts) {
        const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
        const bool has = K::NG == 1 ? grp == 0 : grp < K::PP;
        if (!has) return;
        BTK_UNROLL
        for (int q = 0; q < (K::NG == 1 ? K::PP : 1); q++) {
          const int pp = K::NG == 1 ? q : grp;
          const int P = warp * K::PP + pp, dst_rank = P % csz;
          if (dst_rank == rank) continue;
          float4* dst = reinterpret_cast<float4*>(rbuf + ((long long)rank * own + P / csz) * PW) + gl;
          BTK_UNROLL
          for (int r2 = 0; r2 < G::V / 2; r2++) {
            const cf a = ts.g[q * G::V + 2 * r2], b = ts.g[q * G::V + 2 * r2 + 1];
            float4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
            ctx.cl_send16(dst + r2 * G::L, dst_rank, v, 0);
          }
        }
      });
      ctx.cl_wait(0);                                       // the partials of the pairs this rank owns have landed
      ctx.par([&](int tid, TS& ts) {
        const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
        const bool has = K::NG == 1 ? grp == 0 : grp < K::PP;
        if (!has) return;
        BTK_UNROLL
        for (int q = 0; q < (K::NG == 1 ? K::PP : 1); q++) {
          const int pp = K::NG == 1 ? q : grp;
          const int P = warp * K::PP + pp;
          if (P % csz != rank) continue;
          for (int src = 0; src < csz; src++) {
            if (src == rank) continue;
            const float4* sp = reinterpret_cast<const float4*>(rbuf + ((long long)src * own + P / csz) * PW) + gl;
            BTK_UNROLL
            for (int r2 = 0; r2 < G::V / 2; r2++) {
              const float4 v = sp[r2 * G::L];
              ts.g[q * G::V + 2 * r2] = cadd(ts.g[q * G::V + 2 * r2], mk(v.x, v.y));
              ts.g[q * G::V + 2 * r2 + 1] = cadd(ts.g[q * G::V + 2 * r2 + 1], mk(v.z, v.w));
            }
          }
        }
      });
      ctx.sync();                                           // the receive buffer is the exchange buffer of the transforms below
    } else if (S.valias) {
      ctx.sync();                                           // every warp is past its last read of the aliased window
    }
    synth_transform_store<K>(ctx, s_xbuf, s_twa, s_twb, s_vcur, tau_base, csz, rank);
    if (csz > 1) {
      // ---- only the pairs this rank owns (complete sums) were transformed; their v frames go to every other rank (the
      // local copy is already in place)
      ctx.sync();
      ctx.par([&](int tid, TS&) {
        for (int P = rank; P < K::NW * K::PP; P += csz) {
          float4* loc = reinterpret_cast<float4*>(s_vcur + (long long)(2 * P) * M_);
          for (int i = tid; i < 2 * M_ / 4; i += K::NT) {
            const float4 v = loc[i];
            for (int o = 1; o < csz; o++) ctx.cl_send16(loc + i, (rank + o) % csz, v, 1);
          }
        }
      });
      ctx.cl_wait(1);                                       // the v frames of the other ranks' pairs have landed
    } else {
      ctx.sync();
    }
    {
      const int dn = K::D / csz;
      synth_emit<K>(ctx, L, p.taps_g, s_vhist, s_vcur, out, m, p.pd_s, p.gain, tau_base, wk.j0, wk.nj, rank * dn, dn);
    }
    ctx.sync();
    synth_roll_history<K>(ctx, L, s_vhist, s_vcur);
    ctx.sync();
    if (S.valias) ctx.release(s_last);
  }
}

}  // namespace btk
