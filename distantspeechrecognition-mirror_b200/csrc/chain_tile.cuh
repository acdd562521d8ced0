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
// The tile program is written against a context (par / sync / syncwarp) so that the CPU-only
// tests can execute the very same code sequentially (tests/emu); the library itself only
// instantiates the device context.
#pragma once

#include "fb_core.cuh"

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
  const float* taps_h;   // [N] analysis prototype
  const float* taps_g;   // [m][M]  gp[k][q] = g[M-1-q + M k]
  const cf* wts;         // [Cpad][M] Hermitian-extended conj weights (zero rows for c >= C)
  const cf* tw;          // [M] e^{+j 2 pi t / M}
  int C, Cpad;
  int m;                 // prototype length factor
  int pd_s;              // synthesis processing delay (frames)
  int laN;               // analysis look-ahead (frames skipped)
  int gain;              // synthesis gainFactor
};

template <int M_, int R_>
struct ChainCfg {
  typedef FFTGeom<M_> G;
  static constexpr int M = M_, R = R_, D = M_ / R_;
  static constexpr int NW = (M_ >= 1024) ? 4 : 8;   // warps per CTA, one frame PAIR per warp per iteration
                                                    // (M = 1024 would not fit 227 KB of shared memory with 8)
  static constexpr int NT = NW * 32;
  static constexpr int W = 2 * NW;             // analysis frames per iteration
  static constexpr int CG = 4;                 // channels staged per pass (one float4 per time step)
  static constexpr int NG = G::NG;             // lane groups per warp = channels processed concurrently
  static constexpr int E = G::Ra / R_;         // registers between members of one residue class
  static_assert(G::Ra % R_ == 0, "decimation factor must divide the first radix");
  static_assert(CG % NG == 0, "channel group must be a multiple of the lane groups per warp");
};

struct ChainSmem {
  // offsets in BYTES from the dynamic shared memory base
  int tw, taps, xs, xbuf, vbuf, total;
  int RL;      // row length (floats) of one staged channel
  int win;     // samples per staged window
  int VR;      // slots in the v ring
};

template <int M_, int R_>
BTK_HD ChainSmem chain_smem_layout(int m) {
  typedef ChainCfg<M_, R_> K;
  ChainSmem s;
  const int N = M_ * m;
  s.win = (K::W - 1) * K::D + N;
  int rl = (s.win + 31) & ~31;
  if (K::NG > 1) rl += 32 / K::NG;             // rows of concurrently-read channels land in disjoint banks
  s.RL = rl;
  s.VR = K::W + m * R_ - 1;
  int off = 0;
  s.tw = off;   off += M_ * 8;
  s.taps = off; off += N * 4;
  s.xs = off;   off += K::CG * rl * 4;
  off = (off + 15) & ~15;
  s.xbuf = off; off += K::NW * K::NG * K::G::XBUF * 8;
  s.vbuf = off; off += s.VR * M_ * 4;
  s.total = off;
  return s;
}

template <int M_> struct ChainThreadState {
  cf z[FFTGeom<M_>::V];
  cf g[FFTGeom<M_>::V];
};

// ---------------------------------------------------------------------------------------------
// Polyphase windowing of one frame pair of one staged channel (modulated.cc:419-434), using the
// overlap between the two frames: for the residue class rho (mod D)
//     s_t = x[(i1+1) D - 1 - rho - D t],  h_t = h[rho + D t],  t = a + R k
//     u_{i1}[rho + D a] = sum_k h_t s_t ,   u_{i0}[rho + D a] = sum_k h_t s_{t+1}
// i.e. m R + 1 sample loads and m R tap loads feed 2 m R multiply-adds.
// z[r].x <- u_{i0}, z[r].y <- u_{i1} in the canonical register layout.
// ---------------------------------------------------------------------------------------------
template <int M_, int R_>
BTK_HD void polyphase_pair(cf* z, int gl, const float* row, int n1, const float* taps, int m) {
  typedef ChainCfg<M_, R_> K;
  typedef typename K::G G;
  BTK_UNROLL
  for (int rep = 0; rep < G::RepA; rep++) {
    BTK_UNROLL
    for (int e0 = 0; e0 < K::E; e0++) {
      const int rho = gl + G::L * rep + G::JA * e0;
      const float* xp = row + (n1 - rho);
      const float* hp = taps + rho;
      float u0[R_], u1[R_];
      BTK_UNROLL
      for (int a = 0; a < R_; a++) { u0[a] = 0.f; u1[a] = 0.f; }
      float hprev = 0.f;
      for (int k = 0; k < m; k++) {
        BTK_UNROLL
        for (int a = 0; a < R_; a++) {
          const int t = a + R_ * k;
          const float s = xp[-K::D * t];
          const float h = hp[K::D * t];
          u1[a] = fmaf(h, s, u1[a]);
          u0[(a + R_ - 1) % R_] = fmaf(hprev, s, u0[(a + R_ - 1) % R_]);
          hprev = h;
        }
      }
      u0[R_ - 1] = fmaf(hprev, xp[-K::D * (R_ * m)], u0[R_ - 1]);
      BTK_UNROLL
      for (int a = 0; a < R_; a++) z[rep * G::Ra + e0 + K::E * a] = mk(u0[a], u1[a]);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Forward transform of the packed pair held in ts.g by lane group 0 of every warp, real parts into
// the v ring:  v_{tau0} + j v_{tau0+1} = FFT_fwd(G)   (modulated.cc:603-607, Re taken implicitly
// because G is the sum of two Hermitian spectra).
// ---------------------------------------------------------------------------------------------
template <int M_, int R_, class Ctx>
BTK_HD void synth_transform_store(Ctx& ctx, const ChainSmem& L, cf* s_xbuf, const cf* s_tw, float* s_v, int it,
                                  int tau_base) {
  typedef ChainCfg<M_, R_> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_> TS;
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
    if (grp == 0) GroupFFT<M_, -1>::step1(ts.g, gl, s_xbuf + (warp * K::NG) * G::XBUF, s_tw);
  });
  ctx.syncwarp();
  if (G::Rb > 1) {
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
      if (grp == 0) GroupFFT<M_, -1>::step2_load(ts.g, gl, s_xbuf + (warp * K::NG) * G::XBUF, s_tw);
    });
    ctx.syncwarp();
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
      if (grp == 0) GroupFFT<M_, -1>::step2_store(ts.g, gl, s_xbuf + (warp * K::NG) * G::XBUF);
    });
    ctx.syncwarp();
  }
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
    if (grp == 0) {
      GroupFFT<M_, -1>::step3(ts.g, gl, s_xbuf + (warp * K::NG) * G::XBUF);
      const int tau0 = tau_base + 2 * warp;
      // v of frames before the stream start is zero (the synthesis buffer starts zeroed, modulated.cc:666-674)
      const float k0 = tau0 >= 0 ? 1.f : 0.f, k1 = tau0 + 1 >= 0 ? 1.f : 0.f;
      const int slot0 = (it * K::W + 2 * warp) % L.VR;
      const int slot1 = slot0 + 1 == L.VR ? 0 : slot0 + 1;
      BTK_UNROLL
      for (int r = 0; r < G::V; r++) {
        const int q = G::index_of(gl, r);
        s_v[slot0 * M_ + q] = ts.g[r].x * k0;
        s_v[slot1 * M_ + q] = ts.g[r].y * k1;
      }
    }
  });
}

// ---------------------------------------------------------------------------------------------
// Polyphase with g + overlap-add (modulated.cc:646-661), one output sample per (frame, d):
//   out_j[D-1-d] = sum_{s<R} w_{j-(R-1-s)}[d + s D],  w_j[q] = sum_k g[M-1-q+M k] v_{j+pd-R k}[q],
//   w_{j'<0} = 0 (the pd priming frames never produce a w: the reference's priming quirk).
// Emits the frames of this iteration that fall inside [j0, j0+nj).
// ---------------------------------------------------------------------------------------------
template <int M_, int R_, class Ctx>
BTK_HD void synth_emit(Ctx& ctx, const ChainSmem& L, const float* taps_g, const float* s_v, float* out, int m,
                       int pd_s, int gain, int it, int tau_base, int j0, int nj) {
  typedef ChainCfg<M_, R_> K;
  typedef ChainThreadState<M_> TS;
  ctx.par([&](int tid, TS&) {
    for (int idx = tid; idx < K::W * K::D; idx += K::NT) {
      const int fo = idx / K::D, d = idx % K::D;
      const int tau = tau_base + fo;
      const int j = tau - pd_s;
      if (j < j0 || j >= j0 + nj) continue;
      float acc = 0.f;
      BTK_UNROLL
      for (int s = 0; s < R_; s++) {
        const int back = R_ - 1 - s;                    // w_{j-back} contributes block s
        if (j - back < 0) continue;
        const int q = d + s * K::D;
        int slot = (it * K::W + fo - back) % L.VR;      // >= 0 for every emitted frame
        if (slot < 0) slot += L.VR;
        float w = 0.f;
        for (int k = 0; k < m; k++) {
          w = fmaf(taps_g[k * M_ + q], s_v[slot * M_ + q], w);
          slot -= R_;
          if (slot < 0) slot += L.VR;
        }
        acc += w;
      }
      if (gain > 0) acc *= (float)gain;
      out[(long long)j * K::D + (K::D - 1 - d)] = acc;
    }
  });
}

// One analysis round of a warp: polyphase of the staged channel + backward transform, leaving
// Z = X_{tau0} + j X_{tau0+1} of channel (round*NG + grp) in ts.z (canonical layout).
template <int M_, int R_, class Ctx>
BTK_HD void analysis_round(Ctx& ctx, const ChainSmem& L, const float* s_xs, const float* s_taps, cf* s_xbuf,
                           const cf* s_tw, int m, int round) {
  typedef ChainCfg<M_, R_> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_> TS;
  const int N = M_ * m;
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
    const int c_local = round * K::NG + grp;
    // newest sample of the second frame of the pair, as an index into the staged row
    const int n1 = (2 * warp + 1) * K::D + N - 1;
    polyphase_pair<M_, R_>(ts.z, gl, s_xs + c_local * L.RL, n1, s_taps, m);
    GroupFFT<M_, +1>::step1(ts.z, gl, s_xbuf + (warp * K::NG + grp) * G::XBUF, s_tw);
  });
  ctx.syncwarp();
  if (G::Rb > 1) {
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
      GroupFFT<M_, +1>::step2_load(ts.z, gl, s_xbuf + (warp * K::NG + grp) * G::XBUF, s_tw);
    });
    ctx.syncwarp();
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
      GroupFFT<M_, +1>::step2_store(ts.z, gl, s_xbuf + (warp * K::NG + grp) * G::XBUF);
    });
    ctx.syncwarp();
  }
  ctx.par([&](int tid, TS& ts) {
    const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
    GroupFFT<M_, +1>::step3(ts.z, gl, s_xbuf + (warp * K::NG + grp) * G::XBUF);
  });
}

// Stage CG channels [cg0, cg0+CG) of the window starting at sample t_lo: [t][C] -> s_xs[c][t].
template <int M_, int R_, class Ctx>
BTK_HD void stage_window(Ctx& ctx, const ChainSmem& L, float* s_xs, const float* pcm, int C, int T, long long t_lo,
                         int cg0, bool vec4) {
  typedef ChainCfg<M_, R_> K;
  typedef ChainThreadState<M_> TS;
  ctx.par([&](int tid, TS&) {
    for (int tt = tid; tt < L.win; tt += K::NT) {
      const long long t = t_lo + tt;
      float x[K::CG];
      BTK_UNROLL
      for (int c = 0; c < K::CG; c++) x[c] = 0.f;
      if (t >= 0 && t < T) {
        const float* src = pcm + t * C + cg0;
        if (vec4 && cg0 + K::CG <= C) {
          const float4 q = *reinterpret_cast<const float4*>(src);
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        } else {
          BTK_UNROLL
          for (int c = 0; c < K::CG; c++) if (cg0 + c < C) x[c] = src[c];
        }
      }
      BTK_UNROLL
      for (int c = 0; c < K::CG; c++) s_xs[c * L.RL + tt] = x[c];
    }
  });
}

// ---------------------------------------------------------------------------------------------
// The tile program.  Ctx provides:
//   template<class F> void par(F f)   run f(tid, ChainThreadState&) for every thread of the CTA
//   void sync()                       CTA barrier;   void syncwarp()   warp barrier
// Every cross-thread shared-memory dependency crosses a par() boundary followed by a barrier.
// ---------------------------------------------------------------------------------------------
template <int M_, int R_, class Ctx>
BTK_HD void chain_tile(Ctx& ctx, const ChainParams& p, unsigned char* smem, int work_id) {
  typedef ChainCfg<M_, R_> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_> TS;
  const int m = p.m;
  const int N = M_ * m;
  const int H = m * R_ - 1;                                 // v history needed before a frame
  const ChainSmem L = chain_smem_layout<M_, R_>(m);
  cf* s_tw = reinterpret_cast<cf*>(smem + L.tw);
  float* s_taps = reinterpret_cast<float*>(smem + L.taps);
  float* s_xs = reinterpret_cast<float*>(smem + L.xs);
  cf* s_xbuf = reinterpret_cast<cf*>(smem + L.xbuf);
  float* s_v = reinterpret_cast<float*>(smem + L.vbuf);

  const WorkItem wk = p.work[work_id];
  const RecDesc rec = p.recs[wk.rec];
  const float* pcm = p.pcm + rec.pcm_off;
  float* out = p.out + rec.out_off;
  const int C = p.C;
  const int a_start = wk.j0 + p.pd_s - H;                   // first analysis frame this chunk computes
  const int n_it = (wk.nj + H + K::W - 1) / K::W;
  const bool vec4 = (C % 4 == 0) && (rec.pcm_off % 4 == 0);

  ctx.par([&](int tid, TS&) {
    for (int i = tid; i < M_; i += K::NT) s_tw[i] = p.tw[i];
    for (int i = tid; i < N; i += K::NT) s_taps[i] = p.taps_h[i];
  });
  ctx.sync();

  for (int it = 0; it < n_it; it++) {
    const int tau_base = a_start + it * K::W;
    // oldest sample of the staged window: frame i = tau_base + laN needs x[(i+1) D - N .. (i+1) D - 1]
    const long long t_lo = (long long)(tau_base + p.laN + 1) * K::D - N;

    ctx.par([&](int, TS& ts) {
      BTK_UNROLL
      for (int r = 0; r < G::V; r++) ts.g[r] = mk(0.f, 0.f);
    });

    for (int cg0 = 0; cg0 < p.Cpad; cg0 += K::CG) {
      stage_window<M_, R_>(ctx, L, s_xs, pcm, C, rec.T, t_lo, cg0, vec4);
      ctx.sync();

      // ---- per warp: frame pair (tau0, tau0+1); per lane group: one channel per round
      for (int round = 0; round < K::CG / K::NG; round++) {
        analysis_round<M_, R_>(ctx, L, s_xs, s_taps, s_xbuf, s_tw, m, round);
        ctx.par([&](int tid, TS& ts) {
          const int lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
          const int c = cg0 + round * K::NG + grp;
          const cf* wrow = p.wts + (long long)c * M_;
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) cfma(ts.g[r], ts.z[r], wrow[G::index_of(gl, r)]);
        });
        ctx.syncwarp();
      }
      ctx.sync();
    }

    // ---- sum the partial G of the lane groups of each warp (channels were split across groups)
    if (K::NG > 1) {
      ctx.par([&](int tid, TS& ts) {
        const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
        if (grp > 0) {
          cf* xb = s_xbuf + (warp * K::NG + grp) * G::XBUF;
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) xb[r * G::L + gl] = ts.g[r];
        }
      });
      ctx.syncwarp();
      ctx.par([&](int tid, TS& ts) {
        const int warp = tid >> 5, lane = tid & 31, grp = lane / G::L, gl = lane % G::L;
        if (grp == 0) {
          for (int o = 1; o < K::NG; o++) {
            const cf* xb = s_xbuf + (warp * K::NG + o) * G::XBUF;
            BTK_UNROLL
            for (int r = 0; r < G::V; r++) ts.g[r] = cadd(ts.g[r], xb[r * G::L + gl]);
          }
        }
      });
      ctx.syncwarp();
    }

    synth_transform_store<M_, R_>(ctx, L, s_xbuf, s_tw, s_v, it, tau_base);
    ctx.sync();
    synth_emit<M_, R_>(ctx, L, p.taps_g, s_v, out, m, p.pd_s, p.gain, it, tau_base, wk.j0, wk.nj);
    ctx.sync();
  }
}

}  // namespace btk
