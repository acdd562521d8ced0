// chain_tile.cuh -- the fused hot path: interleaved PCM -> polyphase analysis -> per-subband
// weight apply over channels (SubbandDS / SubbandMVDR) -> synthesis -> PCM, one CTA per chunk
// of output frames of one recording.
//
// Reference behaviour restated here (paths relative to /root/reference/btk):
//   analysis framing + polyphase + backward FFT   modulated/modulated.cc:412-516
//   snapshot transpose + zdotc per bin + mirror   beamformer/beamformer.cc:82-90, 1137-1200, 2583-2635
//   forward FFT (real part) + polyphase + OLA     modulated/modulated.cc:595-664
//
// Data-parallel restructuring (see DESIGN.md for the derivation):
//   * frames i and i+1 of ONE channel are packed as z = u_i + j u_{i+1} into one complex
//     M-point transform, so Z = X_i + j X_{i+1};
//   * the weight table is the Hermitian extension gam_c[k] = conj(w_c[k]) (k <= M/2),
//     gam_c[M-k] = w_c[k], real at k = 0 and M/2 (the reference's synthesis keeps only the real
//     part of the forward FFT, which discards Im Y[0], Im Y[M/2]);  then
//     G[k] = sum_c gam_c[k] Z_c[k] = Y_i[k] + j Y_{i+1}[k] for ALL k with Y Hermitian;
//   * one forward transform of G yields v_i + j v_{i+1} (both real) -- no split/unsplit pass at all.
//
// Shared-memory layouts are chosen for the measured B200 issue rates (tools/ubench):
// every table a lane reads is contiguous per lane (vector LDS with immediate offsets, no index
// arithmetic in the inner loops), the staged sample window is stored residue-major so that the
// m*R+1 samples one lane needs are consecutive floats, and the prototype length factor m is a
// compile-time constant (MT) for the common cases so that the polyphase loops unroll completely.
//
// The tile program is written against a context (par / sync / syncwarp) so that the CPU-only
// tests can execute the very same code sequentially (tests/emu); the library itself only
// instantiates the device context.
#pragma once

#include "fb_core.cuh"

// Staged-window layout of the warp-specialised chain (chain_ws.cuh, WsCfg::RAW): 1 = the layout of the interleaved input,
// which is what the tensor copy engine delivers; 0 = residue-major per channel (register transpose in the producer warps).
#ifndef BTK_WS_RAW
#define BTK_WS_RAW 1
#endif

namespace btk {

struct RecDesc {
  long long pcm_off;   // element offset of this recording's [T][C] block inside pcm
  long long out_off;   // element offset of this recording's output (nblk*D floats)
  int T;               // samples per channel
  int nblk;            // ceil(T / D) output frames
};

struct WorkItem {
  int rec;   // index into recs
  int j0;    // first output frame of the chunk
  int nj;    // number of output frames
};

struct ChainParams {
  const float* pcm;
  float* out;
  const RecDesc* recs;
  const WorkItem* work;
  const float* taps_h;   // [D][TS]   taps_h[rho*TS + t] = h[rho + D t]            (host_tables.h)
  const float* taps_g;   // [m][M]    gp[k][q] = g[M-1-q + M k]
  const cf* wts;         // [Cpad][V/2][L][2] Hermitian-extended conj weights in register order
  long long wts_stride;  // elements between the tables of consecutive recordings (0: one table for the whole batch)
  int no_prefetch;       // skip the L2 prefetch of the next window (many channels: it would evict the rows in use)
  int one_cta;           // launch hint: keep one CTA per SM (host side only, see kern_fb.cuh / capi.cu)
  const cf* twa;         // pass-A twiddles, lane-contiguous (FFTTables)
  const cf* twb;         // pass-B twiddles (three-pass transforms only)
  int C, Cpad;
  int m;                 // prototype length factor
  int pd_s;              // synthesis processing delay (frames)
  int laN;               // analysis look-ahead (frames skipped)
  int gain;              // synthesis gainFactor
  int cluster;           // warp-specialised chain only (chain_ws.cuh): CTAs per work item, the channel groups are split
                         // across the CTAs of a thread-block cluster (0 / 1 = no cluster)
  const void* tmaps;     // warp-specialised chain only: one 128-byte tensor map (CUtensorMap) per recording describing its
                         // interleaved PCM as a [T][C] tensor with boxes of (4 channels x tma_rows time steps); NULL = the
                         // producer warps load the window with 16-byte register loads
  int tma_rows;          // time steps per box (min(D, 256))
  // warp-specialised chain, persistent schedule (chain_ws.cuh::WsSegs): the launch covers items [item0, item0 + n_items),
  // an item = W consecutive output frames of one recording; item_begin[r] = first item of recording r (prefix sums).
  // Every CTA takes a contiguous share of the items.  NULL = one work item per CTA (p.work).
  const int* item_begin;
  int item_q;            // output frames per item (W)
  int item0, n_items;
  int n_rec;             // recordings of the batch (item_begin has n_rec + 1 entries)
  // optional: first item of every CTA of the launch (cta_n + 1 entries, chosen on the host so that every CTA runs the
  // same number of ITERATIONS, host_tables.h::balance_ctas); NULL = equal item counts
  const int* cta_begin;
  int cta_n;
  int no_syn;            // warp-specialised chain, A/B knobs: bit 0 (BTK_WS_SYN=0) the transform warps keep the synthesis side,
                         // bit 1 (BTK_WS_DUAL=0) one channel per windowing pass instead of two
};

template <int M_, int R_, int MT_ = 0, int PP_ = 1>
struct ChainCfg {
  typedef FFTGeom<M_> G;
  static constexpr int M = M_, R = R_, D = M_ / R_;
  static constexpr int MT = MT_;               // compile-time prototype length factor, 0 = runtime
  static constexpr int PP = PP_;               // frame PAIRS per warp per iteration: with 2, a lane windows four
                                               // consecutive frames from one set of tap / sample / weight loads
  // warps per CTA: 8, or 4 where 8 would not leave room for a second CTA on the SM (two frame pairs per warp) or
  // would not fit 227 KB of shared memory at all (M = 1024).  Two independent CTAs per SM matter more than a
  // longer window: one CTA's staging (global-load latency) overlaps the other's transforms.  (M = 512 with 4 warps
  // and two CTAs measured 11-38 % slower than 8 warps and one CTA: the window halo doubles; tools/ab_run2.sh.)
  // (M = 512 without oversampling, D = 512: the window of 8 warps does not fit next to the exchange buffers either)
#ifndef BTK_NW512_4
#define BTK_NW512_4 0      // tuning: 1 = four-warp CTAs (two per SM) for every M = 512 geometry
#endif
  static constexpr int NW = (M_ >= 1024 || (M_ == 512 && (R_ == 1 || BTK_NW512_4)) || (PP_ == 2 && M_ <= 256)) ? 4 : 8;
  static constexpr int NT = NW * 32;
  static constexpr int FW = 2 * PP_;           // frames per warp per iteration
  static constexpr int LV = (FW % 4 == 0) ? 4 : 2;   // floats per shared-memory access of the staged window (a warp's
                                                     // span starts at block FW*warp: 16- or 8-byte aligned)
  static constexpr int W = FW * NW;            // analysis frames per iteration
  static constexpr int CG = 4;                 // channels staged per pass (one float4 per time step)
  static constexpr bool RAW = false;           // staged window layout: residue-major per channel (see chain_ws.cuh for the raw one)
  static constexpr int NG = G::NG;             // lane groups per warp = channels processed concurrently
  static constexpr int XS = PP_;               // exchange buffers per lane group (one per frame pair)
  static constexpr int E = G::Ra / R_;         // registers between members of one residue class
  // lane <-> (lane group, lane inside the group): groups are runs of L consecutive lanes here (chain_ws.cuh interleaves them)
  static BTK_HD int lane_grp(int lane) { return lane / G::L; }
  static BTK_HD int lane_gl(int lane) { return lane % G::L; }
  static constexpr int XPAD = 0;               // complex words between the exchange buffers of consecutive lane groups
  // emit: frames per thread (register blocking of the synthesis polyphase)
  static constexpr int FPT_RAW = (W * D) / NT;
  static constexpr int FPT = FPT_RAW >= 8 ? 8 : (FPT_RAW >= 4 ? 4 : (FPT_RAW >= 2 ? 2 : 1));
  static_assert(G::Ra % R_ == 0, "decimation factor must divide the first radix");
  static_assert(CG % NG == 0, "channel group must be a multiple of the lane groups per warp");
  static_assert(W % FPT == 0, "frames per thread must divide the iteration");
  static_assert(PP_ == 1 || PP_ == 2, "one or two frame pairs per warp");
};

struct ChainSmem {
  // offsets in BYTES from the dynamic shared memory base
  int taps, twa, twb, xs, wts, xbuf, vhist, total;
  int TS;      // floats per residue row of the tap table (m R padded: conflict-free vector loads)
  int TV;      // vector width (floats) of the tap loads
  int NB;      // D-blocks in the staged window
  int SB;      // floats per residue row of the staged window: LV = 2: NB padded so that SB/2 is odd (conflict-free
               // 8-byte accesses); LV = 4: NB rounded up to a multiple of 4, 16-byte quads XOR-swizzled by row (xs_off)
  int CS;      // floats per staged channel
  int H;       // v history frames kept between iterations (m R - 1)
};

BTK_HD constexpr int tap_vec(int mR) { return (mR % 4 == 0) ? 4 : ((mR % 2 == 0) ? 2 : 1); }
BTK_HD constexpr int tap_stride(int mR) {
  const int tv = tap_vec(mR);
  int ts = mR;
  while (((ts / tv) & 1) == 0) ts += tv;
  return ts;
}

template <int M_, int R_, int PP_ = 1>
BTK_HD constexpr ChainSmem chain_smem_layout(int m) {
  typedef ChainCfg<M_, R_, 0, PP_> K;
  typedef FFTTables<M_> FT;
  ChainSmem s = ChainSmem();
  const int mR = m * R_;
  s.TV = tap_vec(mR);
  s.TS = tap_stride(mR);
  s.NB = K::W - 1 + mR;
  int sb = 0;
  if (K::LV == 4) {
    sb = (s.NB + 3) & ~3;
    if (((sb / 4) & 1) != 0) sb += 4;      // an even number of quads per row: rows r and r+4 share a bank group, the swizzle splits them
  } else {
    sb = (s.NB + 1) & ~1;
    if (((sb / 2) & 1) == 0) sb += 2;
  }
  s.SB = sb;
  int cs = K::D * sb;
  if (K::G::L < 16) cs += (16 - (cs & 31) + 32) & 31;   // two lane groups per LDS.64 phase land in disjoint banks
  s.CS = cs;
  s.H = mR - 1;
  int off = 0;
  s.taps = off; off += K::D * s.TS * 4;          off = (off + 15) & ~15;
  s.twa = off;  off += FT::TWA_WORDS * 8;        off = (off + 15) & ~15;
  s.twb = off;  off += FT::TWB_WORDS * 8;        off = (off + 15) & ~15;
  // staged window; the W current v frames of the synthesis side alias it (the window is dead by then)
  int xs_bytes = K::CG * cs * 4;
  if (xs_bytes < K::W * M_ * 4) xs_bytes = K::W * M_ * 4;
  s.xs = off;   off += xs_bytes;                 off = (off + 15) & ~15;
  s.wts = off;  off += K::CG * M_ * 8;
  s.xbuf = off; off += K::NW * K::NG * PP_ * K::G::XBUF * 8;
  s.vhist = off; off += (s.H > 0 ? s.H : 1) * M_ * 4;
  s.total = off;
  return s;
}

template <int M_, int PP_ = 1> struct ChainThreadState {
  cf z[PP_ * FFTGeom<M_>::V];    // [pp][V]
  cf g[PP_ * FFTGeom<M_>::V];
};

// N consecutive floats, the first one aligned to VEC*4 bytes
template <int N, int VEC>
BTK_HD void load_floats(float* dst, const float* src) {
  if (VEC >= 4) {
    BTK_UNROLL
    for (int i = 0; i + 4 <= N; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
    }
    BTK_UNROLL
    for (int i = N & ~3; i < N; i++) dst[i] = src[i];
  } else if (VEC >= 2) {
    BTK_UNROLL
    for (int i = 0; i + 2 <= N; i += 2) {
      const float2 v = *reinterpret_cast<const float2*>(src + i);
      dst[i] = v.x; dst[i + 1] = v.y;
    }
    if (N & 1) dst[N - 1] = src[N - 1];
  } else {
    BTK_UNROLL
    for (int i = 0; i < N; i++) dst[i] = src[i];
  }
}

// Float offset of block group bg (LV consecutive D-blocks) of residue row res inside one staged channel.
// LV = 4: rows are 24 B * k apart, so rows r and r+4 would hit the same bank group; quad index ^ ((res >> 2) & 1)
// makes any 8 consecutive rows conflict-free for 16-byte accesses without padding the rows.
template <int LV>
BTK_HD int xs_off(int res, int bg, int SB) {
  return LV == 4 ? res * SB + ((bg ^ ((res >> 2) & 1)) << 2) : res * SB + 2 * bg;
}

// ---------------------------------------------------------------------------------------------
// Polyphase windowing of PP consecutive frame pairs of one staged channel (modulated.cc:419-434), using the
// overlap between the frames: for the residue class rho (mod D) and pair pp (frames i0 = 2 pp, i1 = i0 + 1)
//     s_t = x[(i1+1) D - 1 - rho - D t],  h_t = h[rho + D t],  t = a + R k
//     u_{i1}[rho + D a] = sum_k h_t s_t ,   u_{i0}[rho + D a] = sum_k h_t s_{t+1}
// i.e. m R + 2 PP - 1 sample loads and m R tap loads feed 2 PP m R multiply-adds.  In the staged window the
// samples of one residue are consecutive (xrow[(D-1-rho) SB + b], b = block index) and so are the taps.
// z[pp][r].x <- u_{i0}, z[pp][r].y <- u_{i1} in the canonical register layout.
// ---------------------------------------------------------------------------------------------
template <class K>
BTK_HD void polyphase_pairs(cf* z, int gl, const float* xch, int warp, const float* taps, const ChainSmem& L, int m) {
  typedef typename K::G G;
  constexpr int R_ = K::R, PP = K::PP;
  BTK_UNROLL
  for (int rep = 0; rep < G::RepA; rep++) {
    BTK_UNROLL
    for (int e0 = 0; e0 < K::E; e0++) {
      const int rho = gl + G::L * rep + G::JA * e0;
      const int res = K::D - 1 - rho;
      // block b of this warp's span is block FW*warp + b of the window; s_t(pp) = that row[2 pp + mR - t]
      const float* hp = taps + rho * L.TS;
      cf u[PP][R_];          // (u_{i0}, u_{i1})[rho + D a] of pair pp
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++) u[pp][a] = mk(0.f, 0.f);
      }
      if (K::MT > 0) {
        constexpr int mR = (K::MT > 0 ? K::MT : 1) * R_;
        constexpr int NX = mR + 1 + 2 * (PP - 1);          // samples needed; read as float2 (the row is padded to even)
        constexpr int NV = (NX + K::LV - 1) / K::LV;       // vector loads
        float xb[NV * K::LV], h[mR];
        if (K::RAW) {
          // raw window [time step][CG]: sample (block b, residue res) of this channel is one float, D * CG floats per block
          BTK_UNROLL
          for (int i = 0; i < NX; i++) xb[i] = xch[((K::FW * warp + i) * K::D + res) * K::CG];
        } else {
          BTK_UNROLL
          for (int i = 0; i < NV; i++)
            load_floats<K::LV, K::LV>(xb + i * K::LV, xch + xs_off<K::LV>(res, (K::FW / K::LV) * warp + i, L.SB));
        }
        load_floats<mR, tap_vec(mR)>(h, hp);
        // (u_{i0}, u_{i1}) += h_t (s_{t+1}, s_t) = h_t (xb[2pp+mR-t-1], xb[2pp+mR-t]): where that is an aligned pair of the
        // float2 loads (mR-t-1 even) it is one packed FFMA2, otherwise two scalar ones
        BTK_UNROLL
        for (int t = 0; t < mR; t++) {
          BTK_UNROLL
          for (int pp = 0; pp < PP; pp++) {
            const int i0 = 2 * pp + mR - t - 1;
            if ((i0 & 1) == 0) {
              u[pp][t % R_] = cfma_real(mk(xb[i0], xb[i0 + 1]), h[t], u[pp][t % R_]);
            } else {
              u[pp][t % R_].y = fmaf(h[t], xb[i0 + 1], u[pp][t % R_].y);
              u[pp][t % R_].x = fmaf(h[t], xb[i0], u[pp][t % R_].x);
            }
          }
        }
      } else {
        const int mR = m * R_;
        auto xs_at = [&](int b) {        // block b of this warp's span
          const int blk = K::FW * warp + b;
          if (K::RAW) return xch[(blk * K::D + res) * K::CG];
          return xch[xs_off<K::LV>(res, blk / K::LV, L.SB) + blk % K::LV];
        };
        for (int k = 0; k < m; k++) {
          BTK_UNROLL
          for (int a = 0; a < R_; a++) {
            const int t = a + R_ * k;
            const float h = hp[t];
            BTK_UNROLL
            for (int pp = 0; pp < PP; pp++) {
              u[pp][a].y = fmaf(h, xs_at(2 * pp + mR - t), u[pp][a].y);
              u[pp][a].x = fmaf(h, xs_at(2 * pp + mR - t - 1), u[pp][a].x);
            }
          }
        }
      }
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++) z[pp * G::V + rep * G::Ra + e0 + K::E * a] = u[pp][a];
      }
    }
  }
}

// The same windowing for TWO channels at once out of the raw window (chain_ws.cuh, [time step][CG]): the samples of
// channels (A, B) = (2 grp, 2 grp + 1) of the stage are neighbours, so ONE 8-byte load brings both -- a warp-wide load then
// uses every byte of the 16-byte granules it touches (two lane groups x two channels) where the one-channel loads of
// polyphase_pairs use half of them and pay a 2-way bank conflict for it; the taps are loaded once for both channels.
// Channel A's windowed frames go to z like in polyphase_pairs; channel B's are handed, residue by residue, to
// park(step, vals) -- 2 PP R floats: (pp, a) -> vals[(pp R + a) 2 + {0, 1}] = (u_{i0}, u_{i1})[rho + D a] -- and come back
// as a whole through polyphase_unpark's `fetch` when their round starts.  Compile-time prototype length only.
template <class K, class Park>
BTK_HD void polyphase_pairs2(cf* z, int gl, const float* xch2, int warp, const float* taps, const ChainSmem& L, Park park) {
  typedef typename K::G G;
  constexpr int R_ = K::R, PP = K::PP;
  constexpr int mR = (K::MT > 0 ? K::MT : 1) * R_;
  constexpr int NX = mR + 1 + 2 * (PP - 1);
  static_assert(K::MT > 0 && K::CG % 2 == 0, "compile-time prototype length, channel pairs");
  BTK_UNROLL
  for (int rep = 0; rep < G::RepA; rep++) {
    BTK_UNROLL
    for (int e0 = 0; e0 < K::E; e0++) {
      const int rho = gl + G::L * rep + G::JA * e0;
      const int res = K::D - 1 - rho;
      const float* hp = taps + rho * L.TS;
      float xa[NX], xb[NX], h[mR];
      BTK_UNROLL
      for (int i = 0; i < NX; i++) {
        const float2 v = *reinterpret_cast<const float2*>(xch2 + ((K::FW * warp + i) * K::D + res) * K::CG);
        xa[i] = v.x; xb[i] = v.y;
      }
      load_floats<mR, tap_vec(mR)>(h, hp);
      cf ua[PP][R_], ub[PP][R_];
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++) { ua[pp][a] = mk(0.f, 0.f); ub[pp][a] = mk(0.f, 0.f); }
      }
      BTK_UNROLL
      for (int t = 0; t < mR; t++) {
        BTK_UNROLL
        for (int pp = 0; pp < PP; pp++) {
          const int i0 = 2 * pp + mR - t - 1;
          if ((i0 & 1) == 0) {
            ua[pp][t % R_] = cfma_real(mk(xa[i0], xa[i0 + 1]), h[t], ua[pp][t % R_]);
            ub[pp][t % R_] = cfma_real(mk(xb[i0], xb[i0 + 1]), h[t], ub[pp][t % R_]);
          } else {
            ua[pp][t % R_].y = fmaf(h[t], xa[i0 + 1], ua[pp][t % R_].y);
            ua[pp][t % R_].x = fmaf(h[t], xa[i0], ua[pp][t % R_].x);
            ub[pp][t % R_].y = fmaf(h[t], xb[i0 + 1], ub[pp][t % R_].y);
            ub[pp][t % R_].x = fmaf(h[t], xb[i0], ub[pp][t % R_].x);
          }
        }
      }
      float pk[2 * PP * R_];
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++) {
          z[pp * G::V + rep * G::Ra + e0 + K::E * a] = ua[pp][a];
          pk[(pp * R_ + a) * 2] = ub[pp][a].x;
          pk[(pp * R_ + a) * 2 + 1] = ub[pp][a].y;
        }
      }
      park(rep * K::E + e0, pk);
    }
  }
}

// Channel B's windowed frames back into z: fetch(col0, vals) delivers PC consecutive parked floats (step-major, see above).
template <class K, class Fetch>
BTK_HD void polyphase_unpark(cf* z, Fetch fetch) {
  typedef typename K::G G;
  constexpr int R_ = K::R, PP = K::PP, NPK = 2 * PP * R_, NSTEP = G::RepA * K::E, TOT = NPK * NSTEP;
  constexpr int PC = TOT < 32 ? TOT : 32;                       // floats per fetch
  static_assert(PC % NPK == 0 && TOT % PC == 0, "whole steps per fetch");
  BTK_UNROLL
  for (int c0 = 0; c0 < TOT; c0 += PC) {
    float vals[PC];
    fetch(c0, vals);
    BTK_UNROLL
    for (int s = 0; s < PC / NPK; s++) {
      const int step = c0 / NPK + s, rep = step / K::E, e0 = step % K::E;
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++)
          z[pp * G::V + rep * G::Ra + e0 + K::E * a] = mk(vals[s * NPK + (pp * R_ + a) * 2], vals[s * NPK + (pp * R_ + a) * 2 + 1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward transforms of the packed pairs of every warp, real parts into the current-v area:
//   v_{tau0} + j v_{tau0+1} = FFT_fwd(G)   (modulated.cc:603-607, Re taken implicitly because G is the sum of
// two Hermitian spectra).  With one lane group per warp (NG == 1) group 0 transforms its PP pairs one after the
// other; otherwise group pp (< PP) transforms pair pp, which synth_gather_pairs left in its ts.g[0..V).
// ---------------------------------------------------------------------------------------------
// (pair_mod, pair_rem): only the frame pairs P = warp * PP + pp with P % pair_mod == pair_rem are transformed and stored
// (cluster mode of chain_ws.cuh: the other pairs' sums live in other CTAs); every pair by default.  A lane that owns
// several pairs (NG == 1 with PP > 1) is selected by its first pair; no configuration combines that with a cluster.
template <class K, class Ctx>
BTK_HD void synth_transform_store(Ctx& ctx, cf* s_xbuf, const cf* s_twa, const cf* s_twb, float* s_vcur, int tau_base,
                                  int pair_mod = 1, int pair_rem = 0) {
  typedef typename K::G G;
  constexpr int M_ = K::M;
  constexpr int NP = K::NG == 1 ? K::PP : 1;                 // transforms per owning lane
  typedef ChainThreadState<M_, K::PP> TS;
  auto owns = [&](int warp, int grp) {
    if (!(K::NG == 1 ? grp == 0 : grp < K::PP)) return false;
    return pair_mod <= 1 || (warp * K::PP + (K::NG == 1 ? 0 : grp)) % pair_mod == pair_rem;
  };
  static_assert(K::NG > 1 || K::XS == K::PP, "a lane owning PP transforms needs PP exchange buffers");
  auto slot = [&](int warp, int grp) { return s_xbuf + ((warp * K::NG + grp) * K::XS) * G::XBUF + (warp * K::NG + grp) * K::XPAD; };
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    if (owns(warp, grp)) {
      if (G::ASYM) GroupFFT<M_, -1>::template inv_step1_multi<NP>(ts.g, gl, slot(warp, grp), s_twa);
      else GroupFFT<M_, -1>::template step1_multi<NP>(ts.g, gl, slot(warp, grp), s_twa);
    }
  });
  ctx.syncwarp();
  if (G::Rb > 1) {
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      if (owns(warp, grp)) {
        BTK_UNROLL
        for (int pp = 0; pp < NP; pp++) GroupFFT<M_, -1>::step2_load(ts.g + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF, s_twb);
      }
    });
    ctx.syncwarp();
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      if (owns(warp, grp)) {
        BTK_UNROLL
        for (int pp = 0; pp < NP; pp++) GroupFFT<M_, -1>::step2_store(ts.g + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF);
      }
    });
    ctx.syncwarp();
  }
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    if (owns(warp, grp)) {
      BTK_UNROLL
      for (int pp = 0; pp < NP; pp++) {
        if (G::ASYM) GroupFFT<M_, -1>::inv_step3(ts.g + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF);
        else GroupFFT<M_, -1>::step3(ts.g + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF);
        const int f = K::FW * warp + 2 * (K::NG == 1 ? pp : grp);        // frame of the iteration
        const int tau0 = tau_base + f;
        // v of frames before the stream start is zero (the synthesis buffer starts zeroed, modulated.cc:666-674)
        const float k0 = tau0 >= 0 ? 1.f : 0.f, k1 = tau0 + 1 >= 0 ? 1.f : 0.f;
        if (s_vcur) {
          float* v0 = s_vcur + f * M_;
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) {
            const int q = G::index_of(gl, r);
            v0[q] = ts.g[pp * G::V + r].x * k0;
            v0[M_ + q] = ts.g[pp * G::V + r].y * k1;
          }
        } else {
          // the caller parks the frames itself (chain_ws.cuh: tensor memory): leave them scaled in the registers
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) ts.g[pp * G::V + r] = mk(ts.g[pp * G::V + r].x * k0, ts.g[pp * G::V + r].y * k1);
        }
      }
    }
  });
}

// Sum the partial G of the lane groups of each warp (channels were split across groups) and hand pair pp to
// group pp: afterwards group pp (< PP) holds the complete G of pair pp in ts.g[0..V).  Only for NG > 1.
template <class K, class Ctx>
BTK_HD void synth_gather_pairs(Ctx& ctx, cf* s_xbuf) {
  typedef typename K::G G;
  typedef ChainThreadState<K::M, K::PP> TS;
  static_assert(K::NG == 1 || K::NG >= K::PP, "a lane group per frame pair");
  if (K::NG == 1) return;
  // with one exchange buffer per lane group (XS == 1 < PP) every group parks exactly one partial: needs NG == PP == 2
  static_assert(K::XS == K::PP || (K::XS == 1 && K::NG == 2 && K::PP == 2), "exchange buffers per lane group");
  auto slot = [&](int warp, int grp, int pp) {
    return s_xbuf + ((warp * K::NG + grp) * K::XS + (K::XS == K::PP ? pp : 0)) * G::XBUF + (warp * K::NG + grp) * K::XPAD;
  };
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    BTK_UNROLL
    for (int pp = 0; pp < K::PP; pp++) {
      if (grp != pp) {
        cf* xb = slot(warp, grp, pp);
        BTK_UNROLL
        for (int r = 0; r < G::V; r++) xb[r * G::L + gl] = ts.g[pp * G::V + r];
      }
    }
  });
  ctx.syncwarp();
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    if (grp < K::PP) {
      cf acc[G::V];
      BTK_UNROLL
      for (int r = 0; r < G::V; r++) {
        acc[r] = ts.g[r];
        BTK_UNROLL
        for (int pp = 1; pp < K::PP; pp++) if (grp == pp) acc[r] = ts.g[pp * G::V + r];
      }
      for (int o = 0; o < K::NG; o++) {
        if (o == grp) continue;
        const cf* xb = slot(warp, o, grp);
        BTK_UNROLL
        for (int r = 0; r < G::V; r++) acc[r] = cadd(acc[r], xb[r * G::L + gl]);
      }
      BTK_UNROLL
      for (int r = 0; r < G::V; r++) ts.g[r] = acc[r];
    }
  });
  ctx.syncwarp();
}

// v frame f of the iteration (f in [-H, W)): history frames live in s_vhist, current ones in s_vcur
BTK_HD const float* v_frame(const float* s_vhist, const float* s_vcur, int H, int M, int f) {
  return f >= 0 ? s_vcur + f * M : s_vhist + (H + f) * M;
}

// ---------------------------------------------------------------------------------------------
// Polyphase with g + overlap-add (modulated.cc:646-661):
//   out_j[D-1-d] = sum_{s<R} w_{j-(R-1-s)}[d + s D],  w_j[q] = sum_k g[M-1-q+M k] v_{j+pd-R k}[q],
//   w_{j'<0} = 0 (the pd priming frames never produce a w: the reference's priming quirk).
// One thread owns one d and FPT consecutive frames, so every v value and every tap is loaded once per
// thread.  Emits the frames of this iteration that fall inside [j0, j0+nj).
// ---------------------------------------------------------------------------------------------
template <class K, class Ctx>
BTK_HD void synth_emit(Ctx& ctx, const ChainSmem& L, const float* taps_g, const float* s_vhist, const float* s_vcur,
                       float* out, int m, int pd_s, int gain, int tau_base, int j0, int nj, int d_lo = 0, int d_n = K::D) {
  constexpr int M_ = K::M, R_ = K::R, D = K::D, FPT = K::FPT;
  typedef ChainThreadState<M_, K::PP> TS;
  const float gf = gain > 0 ? (float)gain : 1.f;
  ctx.par([&](int tid, TS&) {
    // (d_lo, d_n): the slice of every output frame this CTA writes (the whole frame unless a cluster shares the item)
    for (int u = tid; u < d_n * (K::W / FPT); u += K::NT) {
      const int d = d_lo + u % d_n, f0 = (u / d_n) * FPT;        // frames f0 .. f0+FPT-1 of this iteration
      const int jb = tau_base + f0 - pd_s;            // output frame index of f0
      if (jb + FPT <= j0 || jb >= j0 + nj) continue;
      float acc[FPT];
      BTK_UNROLL
      for (int f = 0; f < FPT; f++) acc[f] = 0.f;
      BTK_UNROLL
      for (int s = 0; s < R_; s++) {
        const int back = R_ - 1 - s;                  // w_{j-back} contributes block s
        const int q = d + s * D;
        if (K::MT > 0) {
          constexpr int mm = K::MT > 0 ? K::MT : 1;
          constexpr int NV = FPT + R_ * (mm - 1);      // v frames f0-back-R(m-1) .. f0-back+FPT-1
          float g[mm], v[NV];
          BTK_UNROLL
          for (int k = 0; k < mm; k++) g[k] = taps_g[k * M_ + q];
          BTK_UNROLL
          for (int i = 0; i < NV; i++) v[i] = v_frame(s_vhist, s_vcur, L.H, M_, f0 - back - R_ * (mm - 1) + i)[q];
#ifndef BTK_EMIT_SCALAR
          if (R_ % 2 == 0 && FPT % 2 == 0) {
            // frames f, f+1 (f even) read v[f + R (mm-1-k)] and its neighbour: with R even that is an aligned register
            // pair for every tap, so the two frames share one packed multiply-add per tap
            BTK_UNROLL
            for (int f = 0; f < FPT; f += 2) {
              cf w = mk(0.f, 0.f);
              BTK_UNROLL
              for (int k = 0; k < mm; k++) {
                const int i0 = f + R_ * (mm - 1) - R_ * k;
                w = cfma_real(mk(v[i0], v[i0 + 1]), g[k], w);
              }
              if (jb + f - back >= 0) acc[f] += w.x;
              if (jb + f + 1 - back >= 0) acc[f + 1] += w.y;
            }
          } else
#endif
          {
            BTK_UNROLL
            for (int f = 0; f < FPT; f++) {
              float w = 0.f;
              BTK_UNROLL
              for (int k = 0; k < mm; k++) w = fmaf(g[k], v[f + R_ * (mm - 1) - R_ * k], w);
              if (jb + f - back >= 0) acc[f] += w;
            }
          }
        } else {
          BTK_UNROLL
          for (int f = 0; f < FPT; f++) {
            if (jb + f - back < 0) continue;
            float w = 0.f;
            for (int k = 0; k < m; k++)
              w = fmaf(taps_g[k * M_ + q], v_frame(s_vhist, s_vcur, L.H, M_, f0 + f - back - R_ * k)[q], w);
            acc[f] += w;
          }
        }
      }
      BTK_UNROLL
      for (int f = 0; f < FPT; f++) {
        const int j = jb + f;
        if (j >= j0 && j < j0 + nj) out[(long long)j * D + (D - 1 - d)] = acc[f] * gf;
      }
    }
  });
}

// the last H current v frames become the history of the next iteration
template <class K, class Ctx>
BTK_HD void synth_roll_history(Ctx& ctx, const ChainSmem& L, float* s_vhist, const float* s_vcur) {
  typedef ChainThreadState<K::M, K::PP> TS;
  // History frame f (f in [-H, 0)) of the next iteration is frame f + W of this one.  With H <= W that is always a
  // current frame (one pass).  With H > W (long prototypes at high decimation: m R - 1 > W) the older part comes
  // from the history itself, shifted down by W frames: copied in ascending slices of W frames with a barrier
  // between slices, so that a slice never reads what the same pass writes.
  const int HM = L.H * K::M, WM = K::W * K::M;
  for (int base = 0; base < HM; base += WM) {
    ctx.par([&](int tid, TS&) {
      const int hi = base + WM < HM ? base + WM : HM;
      for (int i = base + tid; i < hi; i += K::NT) s_vhist[i] = i + WM < HM ? s_vhist[i + WM] : s_vcur[i + WM - HM];
    });
    if (base + WM < HM) ctx.sync();
  }
}

// One analysis round of a warp: polyphase of the staged channel + backward transforms, leaving
// Z = X_{tau0} + j X_{tau0+1} of channel (round*NG + grp), for each of the warp's PP frame pairs, in ts.z[pp][V]
// (canonical layout).
// (fill(z, tid, grp, gl) puts the windowed frame pairs of the lane's channel into z: polyphase_pairs by default, see
// analysis_round below; the two-channel variant of chain_ws.cuh windows or fetches them itself)
template <class K, class Ctx, class Fill>
BTK_HD void analysis_round_fill(Ctx& ctx, cf* s_xbuf, const cf* s_twa, const cf* s_twb, Fill fill) {
  typedef typename K::G G;
  constexpr int M_ = K::M;
  typedef ChainThreadState<M_, K::PP> TS;
  auto slot = [&](int warp, int grp) { return s_xbuf + ((warp * K::NG + grp) * K::XS) * G::XBUF + (warp * K::NG + grp) * K::XPAD; };
  if constexpr (K::XS < K::PP) {
    // the frame pairs of a lane take turns on one exchange buffer: pass A of all pairs (one set of twiddles), then
    // scatter / gather pair by pair, then the final radix passes
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      fill(ts.z, tid, grp, gl);
      GroupFFT<M_, +1>::template step1_twiddle<K::PP>(ts.z, gl, s_twa);
      GroupFFT<M_, +1>::step1_scatter(ts.z, gl, slot(warp, grp));
    });
    ctx.syncwarp();
    for (int pp = 0; pp < K::PP; pp++) {
      ctx.par([&](int tid, TS& ts) {
        const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
        GroupFFT<M_, +1>::step3_gather(ts.z + pp * G::V, gl, slot(warp, grp));
      });
      ctx.syncwarp();
      if (pp + 1 < K::PP) {
        ctx.par([&](int tid, TS& ts) {
          const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
          GroupFFT<M_, +1>::step1_scatter(ts.z + (pp + 1) * G::V, gl, slot(warp, grp));
        });
        ctx.syncwarp();
      }
    }
    ctx.par([&](int, TS& ts) {
      BTK_UNROLL
      for (int pp = 0; pp < K::PP; pp++) GroupFFT<M_, +1>::step3_dft(ts.z + pp * G::V);
    });
    return;
  }
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    fill(ts.z, tid, grp, gl);
    GroupFFT<M_, +1>::template step1_multi<K::PP>(ts.z, gl, slot(warp, grp), s_twa);
  });
  ctx.syncwarp();
  if (G::Rb > 1) {
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      BTK_UNROLL
      for (int pp = 0; pp < K::PP; pp++) GroupFFT<M_, +1>::step2_load(ts.z + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF, s_twb);
    });
    ctx.syncwarp();
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      BTK_UNROLL
      for (int pp = 0; pp < K::PP; pp++) GroupFFT<M_, +1>::step2_store(ts.z + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF);
    });
    ctx.syncwarp();
  }
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
    BTK_UNROLL
    for (int pp = 0; pp < K::PP; pp++) GroupFFT<M_, +1>::step3(ts.z + pp * G::V, gl, slot(warp, grp) + pp * G::XBUF);
  });
}

template <class K, class Ctx>
BTK_HD void analysis_round(Ctx& ctx, const ChainSmem& L, const float* s_xs, const float* s_taps, cf* s_xbuf,
                           const cf* s_twa, const cf* s_twb, int m, int round) {
  analysis_round_fill<K>(ctx, s_xbuf, s_twa, s_twb, [&](cf* z, int tid, int grp, int gl) {
    const int c_local = round * K::NG + grp;
    polyphase_pairs<K>(z, gl, K::RAW ? s_xs + c_local : s_xs + c_local * L.CS, tid >> 5, s_taps, L, m);
  });
}

// Stage CG channels [cg0, cg0+CG) of the window starting at sample t_lo: pcm [t][C] -> s_xs[c][t mod D][t div D]
// (residue-major), and (when wts != 0) the weights of those channels, already in register order, into s_wts.
// One task = one residue x LV consecutive D-blocks: LV coalesced float4 loads (4 channels of one time step each;
// consecutive lanes take consecutive time steps) are transposed in registers into ONE 16- or 8-byte store per
// channel.  All loads of a batch of tasks are issued before its first store (12 in flight per thread).
template <class K, class Ctx>
BTK_HD void stage_window(Ctx& ctx, const ChainSmem& L, float* s_xs, const float* pcm, int C, int T, long long t_lo,
                         int cg0, bool vec4, float4* s_wts, const cf* wts) {
  typedef ChainThreadState<K::M, K::PP> TS;
  constexpr int D = K::D, LV = K::LV;
  static_assert(K::CG == 4, "one float4 per time step");
  constexpr int WPT = (K::CG * K::M / 2 + K::NT - 1) / K::NT;   // weight float4 per thread
#ifndef BTK_STAGE_LOADS
#define BTK_STAGE_LOADS 12
#endif
  constexpr int TB = BTK_STAGE_LOADS / LV;                       // tasks per batch
  ctx.par([&](int tid, TS&) {
    float4 wv[WPT];
    if (wts) {
      const float4* src = reinterpret_cast<const float4*>(wts + (long long)cg0 * K::M);
      BTK_UNROLL
      for (int i = 0; i < WPT; i++) {
        const int idx = tid + i * K::NT;
        if (idx < K::CG * K::M / 2) wv[i] = src[idx];
      }
    }
    // 32-bit sample indices inside the loops (T < 2^31; t may be negative at the stream start).
    const int t0 = (int)t_lo;
    const float* pcm_cg = pcm + cg0;
    const bool v4 = vec4 && cg0 + K::CG <= C;
    const int ntask = D * ((L.NB + LV - 1) / LV);
    // Every channel group walks the same rows of interleaved PCM (a row holds all channels).  Odd groups walk the window
    // backwards: a cyclic walk over a footprint larger than the L2 share of this CTA would miss on every pass, the
    // back-and-forth walk re-reads the most recently touched rows first (many-channel inputs).
    const int nbatch = (ntask + K::NT * TB - 1) / (K::NT * TB);
    const bool backwards = ((cg0 / K::CG) & 1) != 0;
    for (int b = 0; b < nbatch; b++) {
      const int task0 = tid + (backwards ? nbatch - 1 - b : b) * (K::NT * TB);
      float x[TB][LV][K::CG];
      BTK_UNROLL
      for (int k = 0; k < TB; k++) {
        const int task = task0 + k * K::NT;
        const int res = task % D, bg = task / D;
        BTK_UNROLL
        for (int i = 0; i < LV; i++) {
          const int blk = bg * LV + i;
          const int t = t0 + blk * D + res;
          BTK_UNROLL
          for (int c = 0; c < K::CG; c++) x[k][i][c] = 0.f;
          if (task < ntask && blk < L.NB && (unsigned)t < (unsigned)T) {
            const float* src = pcm_cg + (size_t)((unsigned)t) * (unsigned)C;
            if (v4) {
              const float4 q = *reinterpret_cast<const float4*>(src);
              x[k][i][0] = q.x; x[k][i][1] = q.y; x[k][i][2] = q.z; x[k][i][3] = q.w;
            } else {
              BTK_UNROLL
              for (int c = 0; c < K::CG; c++) if (cg0 + c < C) x[k][i][c] = src[c];
            }
          }
        }
      }
      BTK_UNROLL
      for (int k = 0; k < TB; k++) {
        const int task = task0 + k * K::NT;
        if (task < ntask) {
          const int res = task % D, bg = task / D;
          float* dst = s_xs + xs_off<LV>(res, bg, L.SB);
          BTK_UNROLL
          for (int c = 0; c < K::CG; c++) {
            if (LV == 4) {
              float4 v; v.x = x[k][0][c]; v.y = x[k][1][c]; v.z = x[k][LV > 2 ? 2 : 0][c]; v.w = x[k][LV - 1][c];
              *reinterpret_cast<float4*>(dst + c * L.CS) = v;
            } else {
              float2 v; v.x = x[k][0][c]; v.y = x[k][1][c];
              *reinterpret_cast<float2*>(dst + c * L.CS) = v;
            }
          }
        }
      }
    }
    if (wts) {
      BTK_UNROLL
      for (int i = 0; i < WPT; i++) {
        const int idx = tid + i * K::NT;
        if (idx < K::CG * K::M / 2) s_wts[idx] = wv[i];
      }
    }
  });
}

// Ask L2 for the samples the NEXT iteration will stage that this one has not touched: W*D new time steps of
// all C channels.  Issued before the compute rounds so that the next staging pass finds them in L2.
template <class K, class Ctx>
BTK_HD void prefetch_next_window(Ctx& ctx, const float* pcm, int C, int T, long long t_first) {
  typedef ChainThreadState<K::M, K::PP> TS;
  ctx.par([&](int tid, TS&) {
    long long lo = t_first < 0 ? 0 : t_first, hi = t_first + (long long)K::W * K::D;
    if (hi > T) hi = T;
    const long long b0 = lo * C, b1 = hi * C;      // float index range
    for (long long i = b0 + (long long)tid * 32; i < b1; i += (long long)K::NT * 32) BTK_PREFETCH_L2(pcm + i);
  });
}

// tables every tile program needs: residue-major taps and the lane-contiguous twiddles
template <class K, class Ctx>
BTK_HD void load_tables(Ctx& ctx, const ChainSmem& L, unsigned char* smem, const float* taps_h, const cf* twa,
                        const cf* twb) {
  typedef ChainThreadState<K::M, K::PP> TS;
  typedef FFTTables<K::M> FT;
  float* s_taps = reinterpret_cast<float*>(smem + L.taps);
  cf* s_twa = reinterpret_cast<cf*>(smem + L.twa);
  cf* s_twb = reinterpret_cast<cf*>(smem + L.twb);
  // all loads of a batch are issued before its first store: one trip to L2 per batch, not one per element
  // (the element-wise loop cost 2.3 % of the chain kernel's samples at the start of every CTA)
  ctx.par([&](int tid, TS&) {
    constexpr int U = 8;
    const int ntaps = taps_h ? K::D * L.TS : 0;
    float tv[U];
    cf av[U], bv[U];
    BTK_UNROLL
    for (int u = 0; u < U; u++) {
      const int i = tid + u * K::NT;
      if (i < ntaps) tv[u] = taps_h[i];
      if (i < FT::TWA_WORDS) av[u] = twa[i];
      if (i < FT::TWB_WORDS) bv[u] = twb[i];
    }
    BTK_UNROLL
    for (int u = 0; u < U; u++) {
      const int i = tid + u * K::NT;
      if (i < ntaps) s_taps[i] = tv[u];
      if (i < FT::TWA_WORDS) s_twa[i] = av[u];
      if (i < FT::TWB_WORDS) s_twb[i] = bv[u];
    }
    for (int i = tid + U * K::NT; i < ntaps; i += K::NT) s_taps[i] = taps_h[i];
    for (int i = tid + U * K::NT; i < FT::TWA_WORDS; i += K::NT) s_twa[i] = twa[i];
    for (int i = tid + U * K::NT; i < FT::TWB_WORDS; i += K::NT) s_twb[i] = twb[i];
  });
}

// ---------------------------------------------------------------------------------------------
// The tile program.  Ctx provides:
//   template<class F> void par(F f)   run f(tid, ChainThreadState&) for every thread of the CTA
//   void sync()                       CTA barrier;   void syncwarp()   warp barrier
// Every cross-thread shared-memory dependency crosses a par() boundary followed by a barrier.
// ---------------------------------------------------------------------------------------------
template <int M_, int R_, int MT_, int PP_, class Ctx>
BTK_HD void chain_tile(Ctx& ctx, const ChainParams& p, unsigned char* smem, int work_id) {
  typedef ChainCfg<M_, R_, MT_, PP_> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_, PP_> TS;
  const int m = MT_ > 0 ? MT_ : p.m;
  const int N = M_ * m;
  const ChainSmem L = chain_smem_layout<M_, R_, PP_>(m);
  const int H = L.H;                                        // v history needed before a frame
  float* s_taps = reinterpret_cast<float*>(smem + L.taps);
  cf* s_twa = reinterpret_cast<cf*>(smem + L.twa);
  cf* s_twb = reinterpret_cast<cf*>(smem + L.twb);
  float* s_xs = reinterpret_cast<float*>(smem + L.xs);
  float* s_vcur = s_xs;                                     // alias: the window is dead when v is produced
  float4* s_wts = reinterpret_cast<float4*>(smem + L.wts);
  cf* s_xbuf = reinterpret_cast<cf*>(smem + L.xbuf);
  float* s_vhist = reinterpret_cast<float*>(smem + L.vhist);

  const WorkItem wk = p.work[work_id];
  const RecDesc rec = p.recs[wk.rec];
  const float* pcm = p.pcm + rec.pcm_off;
  float* out = p.out + rec.out_off;
  const int C = p.C;
  const int a_start = wk.j0 + p.pd_s - H;                   // first analysis frame this chunk computes
  const int n_it = (wk.nj + H + K::W - 1) / K::W;
  // 16-byte loads need the caller's base pointer aligned as well (an offset device view is legal input: scalar path)
  const bool vec4 = (C % 4 == 0) && (rec.pcm_off % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0);

  load_tables<K>(ctx, L, smem, p.taps_h, p.twa, p.twb);
  ctx.par([&](int tid, TS&) {
    for (int i = tid; i < H * M_; i += K::NT) s_vhist[i] = 0.f;
  });
  ctx.sync();

  for (int it = 0; it < n_it; it++) {
    const int tau_base = a_start + it * K::W;
    // oldest sample of the staged window: frame i = tau_base + laN needs x[(i+1) D - N .. (i+1) D - 1]
    const long long t_lo = (long long)(tau_base + p.laN + 1) * K::D - N;

    ctx.par([&](int, TS& ts) {
      BTK_UNROLL
      for (int r = 0; r < PP_ * G::V; r++) ts.g[r] = mk(0.f, 0.f);
    });

    for (int cg0 = 0; cg0 < p.Cpad; cg0 += K::CG) {
      // samples + the weights of these CG channels (register order: [c][V/2][L] float4 = 2 complex)
      stage_window<K>(ctx, L, s_xs, pcm, C, rec.T, t_lo, cg0, vec4, s_wts, p.wts + (long long)wk.rec * p.wts_stride);
      if (cg0 == 0 && it + 1 < n_it && !p.no_prefetch) prefetch_next_window<K>(ctx, pcm, C, rec.T, t_lo + (long long)L.NB * K::D);
      ctx.sync();

      // ---- per warp: PP frame pairs (tau0, tau0+1); per lane group: one channel per round
      for (int round = 0; round < K::CG / K::NG; round++) {
        analysis_round<K>(ctx, L, s_xs, s_taps, s_xbuf, s_twa, s_twb, m, round);
        ctx.par([&](int tid, TS& ts) {
          const int lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
          const float4* w4 = s_wts + (round * K::NG + grp) * (G::V / 2) * G::L + gl;
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
      ctx.sync();
    }

    synth_gather_pairs<K>(ctx, s_xbuf);
    synth_transform_store<K>(ctx, s_xbuf, s_twa, s_twb, s_vcur, tau_base);
    ctx.sync();
    synth_emit<K>(ctx, L, p.taps_g, s_vhist, s_vcur, out, m, p.pd_s, p.gain, tau_base, wk.j0, wk.nj);
    ctx.sync();
    synth_roll_history<K>(ctx, L, s_vhist, s_vcur);
    ctx.sync();
  }
}

}  // namespace btk
