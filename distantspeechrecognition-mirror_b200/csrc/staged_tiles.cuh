// staged_tiles.cuh -- the reference's stage boundaries materialised: analysis bank -> snapshots,
// and beamformer output -> synthesis bank, built from the same pieces as the fused chain
// (chain_tile.cuh).  Used when a caller asks for the subband snapshots / beamformer outputs
// themselves (the drop-in stream nodes, covariance estimation, the parity tests on X and Y).
//
//   analysis_tile : pcm [T][C]  -> snapshots [F][B][C] complex64, channel innermost -- the layout
//                   SnapShotArray::update produces (reference beamformer/beamformer.cc:82-90)
//   synthesis_tile: Y [F][B] complex64 (bins 0..M/2; the upper half is the conjugate mirror the
//                   reference's beamformers write, beamformer.cc:1189-1194) -> float PCM
#pragma once

#include "chain_tile.cuh"

namespace btk {

struct AnalysisParams {
  const float* pcm;
  cf* snap;               // [F][B][C] per recording at recs[].out_off (complex elements)
  const RecDesc* recs;    // nblk field = number of emitted analysis frames F
  const WorkItem* work;   // j0/nj = first emitted frame / count (multiple of W except the last)
  const float* taps_h;    // residue-major taps (host_tables.h::build_analysis_taps)
  const cf* twa;
  const cf* twb;
  int C, Cpad, m, laN;
  int cg_slices;          // CTAs per work item: slice s takes channel groups s, s + cg_slices, ...
};

template <int M_, int R_, int MT_, class Ctx>
BTK_HD void analysis_tile(Ctx& ctx, const AnalysisParams& p, unsigned char* smem, int work_id) {
  typedef ChainCfg<M_, R_, MT_, 1> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_> TS;
  const int m = MT_ > 0 ? MT_ : p.m;
  const int N = M_ * m, B = M_ / 2 + 1;
  const ChainSmem L = chain_smem_layout<M_, R_>(m);
  float* s_taps = reinterpret_cast<float*>(smem + L.taps);
  cf* s_twa = reinterpret_cast<cf*>(smem + L.twa);
  cf* s_twb = reinterpret_cast<cf*>(smem + L.twb);
  float* s_xs = reinterpret_cast<float*>(smem + L.xs);
  cf* s_xbuf = reinterpret_cast<cf*>(smem + L.xbuf);

  const int slices = p.cg_slices > 0 ? p.cg_slices : 1;
  const WorkItem wk = p.work[work_id / slices];
  const int slice = work_id % slices;
  const RecDesc rec = p.recs[wk.rec];
  const float* pcm = p.pcm + rec.pcm_off;
  cf* snap = p.snap + rec.out_off;
  const int C = p.C, F = rec.nblk;
  const int n_it = (wk.nj + K::W - 1) / K::W;
  const bool vec4 = (C % 4 == 0) && (rec.pcm_off % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0);

  load_tables<K>(ctx, L, smem, p.taps_h, p.twa, p.twb);
  ctx.sync();

  for (int it = 0; it < n_it; it++) {
    const int f_base = wk.j0 + it * K::W;          // emitted frame index t; internal frame i = t + laN
    const long long t_lo = (long long)(f_base + p.laN + 1) * K::D - N;
    for (int cg0 = slice * K::CG; cg0 < p.Cpad; cg0 += slices * K::CG) {
      stage_window<K>(ctx, L, s_xs, pcm, C, rec.T, t_lo, cg0, vec4, (float4*)0, (const cf*)0);
      ctx.sync();
      for (int round = 0; round < K::CG / K::NG; round++) {
        analysis_round<K>(ctx, L, s_xs, s_taps, s_xbuf, s_twa, s_twb, m, round);
        ctx.syncwarp();
        // Z = X_f0 + j X_f1 in natural order through the exchange buffer, so every lane can reach Z[M-k]
        ctx.par([&](int tid, TS& ts) {
          const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
          cf* xb = s_xbuf + (warp * K::NG + grp) * G::XBUF;
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) xb[G::index_of_spec(gl, r)] = ts.z[r];
        });
        ctx.syncwarp();
        ctx.par([&](int tid, TS& ts) {
          const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
          const cf* xb = s_xbuf + (warp * K::NG + grp) * G::XBUF;
          const int c = cg0 + round * K::NG + grp;
          const int f0 = f_base + 2 * warp, f1 = f0 + 1;
          const bool ok0 = c < C && f0 < wk.j0 + wk.nj && f0 < F, ok1 = c < C && f1 < wk.j0 + wk.nj && f1 < F;
#if defined(__CUDA_ARCH__)
          // Two lane groups = two NEIGHBOURING channels of the same bins: the groups swap one frame each (lane ^ L), so that
          // group 0 stores (c, c+1) of frame f0 and group 1 (c-1, c) of frame f1 as ONE 16-byte word -- half the store
          // instructions and half the sector writes of the 8-byte stores below (each of which lands in a sector of its own:
          // consecutive lanes are consecutive BINS, C channels apart)
          const bool pair16 = K::NG == 2 && (C & 1) == 0 && (cg0 + K::CG <= C) &&
                              ((reinterpret_cast<uintptr_t>(snap) & 15) == 0);      // uniform over the warp
          if (pair16) {
            const bool both0 = f0 < wk.j0 + wk.nj && f0 < F, both1 = f1 < wk.j0 + wk.nj && f1 < F;
            BTK_UNROLL
            for (int r = 0; r < G::V; r++) {
              const int k = G::index_of_spec(gl, r);
              const cf zk = ts.z[r];
              const cf zm = xb[(M_ - k) & (M_ - 1)];
              const cf x0 = mk(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)), x1 = mk(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
              const cf give = grp == 0 ? x1 : x0;
              cf got;
              got.x = __shfl_xor_sync(0xffffffffu, give.x, G::L);
              got.y = __shfl_xor_sync(0xffffffffu, give.y, G::L);
              if (k > M_ / 2) continue;
              if (grp == 0) {
                if (both0) *reinterpret_cast<float4*>(snap + ((long long)f0 * B + k) * C + c) = make_float4(x0.x, x0.y, got.x, got.y);
              } else {
                if (both1) *reinterpret_cast<float4*>(snap + ((long long)f1 * B + k) * C + c - 1) = make_float4(got.x, got.y, x1.x, x1.y);
              }
            }
            return;
          }
#endif
          BTK_UNROLL
          for (int r = 0; r < G::V; r++) {
            const int k = G::index_of_spec(gl, r);
            if (k > M_ / 2) continue;
            const cf zk = ts.z[r];
            const cf zm = xb[(M_ - k) & (M_ - 1)];
            // X_f0[k] = (Z[k] + conj Z[M-k]) / 2 ;  X_f1[k] = (Z[k] - conj Z[M-k]) / (2j)
            if (ok0) snap[((long long)f0 * B + k) * C + c] = mk(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            if (ok1) snap[((long long)f1 * B + k) * C + c] = mk(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
          }
        });
        ctx.syncwarp();
      }
      ctx.sync();
    }
  }
}

struct SynthesisParams {
  const cf* Y;            // [F][B] per recording at recs[].pcm_off (complex elements)
  float* out;
  const RecDesc* recs;    // T field = F (frames available), nblk = number of output frames
  const WorkItem* work;
  const float* taps_g;
  const cf* twa;
  const cf* twb;
  int m, pd_s, gain;
};

template <int M_, int R_, int MT_, class Ctx>
BTK_HD void synthesis_tile(Ctx& ctx, const SynthesisParams& p, unsigned char* smem, int work_id) {
  typedef ChainCfg<M_, R_, MT_, 1> K;
  typedef typename K::G G;
  typedef ChainThreadState<M_> TS;
  const int m = MT_ > 0 ? MT_ : p.m;
  const int B = M_ / 2 + 1;
  const ChainSmem L = chain_smem_layout<M_, R_>(m);
  const int H = L.H;
  cf* s_twa = reinterpret_cast<cf*>(smem + L.twa);
  cf* s_twb = reinterpret_cast<cf*>(smem + L.twb);
  cf* s_xbuf = reinterpret_cast<cf*>(smem + L.xbuf);
  float* s_vcur = reinterpret_cast<float*>(smem + L.xs);
  float* s_vhist = reinterpret_cast<float*>(smem + L.vhist);

  const WorkItem wk = p.work[work_id];
  const RecDesc rec = p.recs[wk.rec];
  const cf* Y = p.Y + rec.pcm_off;
  float* out = p.out + rec.out_off;
  const int F = rec.T;
  const int a_start = wk.j0 + p.pd_s - H;
  const int n_it = (wk.nj + H + K::W - 1) / K::W;

  load_tables<K>(ctx, L, smem, (const float*)0, p.twa, p.twb);
  ctx.par([&](int tid, TS&) {
    for (int i = tid; i < H * M_; i += K::NT) s_vhist[i] = 0.f;
  });
  ctx.sync();

  for (int it = 0; it < n_it; it++) {
    const int tau_base = a_start + it * K::W;
    // G = Y_tau0 + j Y_tau1 over all M bins, Hermitian-extended, real at k = 0 and M/2
    ctx.par([&](int tid, TS& ts) {
      const int warp = tid >> 5, lane = tid & 31, grp = K::lane_grp(lane), gl = K::lane_gl(lane);
      if (grp != 0) return;
      const int tau0 = tau_base + 2 * warp, tau1 = tau0 + 1;
      const bool ok0 = tau0 >= 0 && tau0 < F, ok1 = tau1 >= 0 && tau1 < F;
      BTK_UNROLL
      for (int r = 0; r < G::V; r++) {
        const int k = G::index_of_spec(gl, r);
        const int kk = k <= M_ / 2 ? k : M_ - k;
        cf a = ok0 ? Y[(long long)tau0 * B + kk] : mk(0.f, 0.f);
        cf b = ok1 ? Y[(long long)tau1 * B + kk] : mk(0.f, 0.f);
        if (k > M_ / 2) { a.y = -a.y; b.y = -b.y; }
        if (k == 0 || k == M_ / 2) { a.y = 0.f; b.y = 0.f; }
        ts.g[r] = mk(a.x - b.y, a.y + b.x);
      }
    });
    synth_transform_store<K>(ctx, s_xbuf, s_twa, s_twb, s_vcur, tau_base);
    ctx.sync();
    synth_emit<K>(ctx, L, p.taps_g, s_vhist, s_vcur, out, m, p.pd_s, p.gain, tau_base, wk.j0, wk.nj);
    ctx.sync();
    synth_roll_history<K>(ctx, L, s_vhist, s_vcur);
    ctx.sync();
  }
}

}  // namespace btk
