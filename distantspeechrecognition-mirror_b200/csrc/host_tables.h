// host_tables.h -- host-side (plain C++) preparation of the constant tables the kernels consume and
// of the reference's setup-time formulas.  No CUDA here; shared by the C-ABI layer (capi.cu) and by the
// CPU-only emulation tests.  Paths cited are relative to /root/reference/btk.
#pragma once

#include <math.h>
#include <complex>
#include <vector>

#include "chain_tile.cuh"

namespace btk {

typedef std::complex<double> zd;

// OverSampledDFTFilterBank ctor (modulated/modulated.cc:262-300): processing delay / look-ahead.
struct BankGeom {
  int M, m, r, dct, R, D, N, B, pd_a, pd_s, laN;
  BankGeom() {}
  BankGeom(int M_, int m_, int r_, int dct_) : M(M_), m(m_), r(r_), dct(dct_) {
    R = 1 << r; D = M / R; N = M * m; B = M / 2 + 1;
    laN = 0;
    switch (dct) {
      case 1: pd_a = pd_s = m * R - 1; break;
      case 2: pd_a = m * R - 1; pd_s = m * R / 2; laN = m * R / 2 - 1; break;
      default: pd_a = pd_s = 2 * m - 1; break;
    }
  }
  // SampleFeature::next (feature/feature.cc:627-641): ceil(T/D) blocks, last zero padded
  int nblk(long long T) const { return (int)((T + D - 1) / D); }
  // modulated.cc:461-516: nblk real + pd padded frames, laN internal frames skipped
  int analysis_frames(long long T) const { return nblk(T) + pd_a - laN; }
  // modulated.cc:626-642: pd_s frames consumed by priming, then one output per input frame
  int synthesis_frames(int F) const { return F > pd_s ? F - pd_s : 0; }
  // output frames of analysis -> beamformer -> synthesis: nblk, except for delayCompensationType 2 with an odd
  // m R, where pd_a - laN - pd_s = 1 and the reference chain emits one more (all-padding) frame
  int chain_frames(long long T) const { return synthesis_frames(analysis_frames(T)); }
};

// Lane-contiguous twiddle tables of the M-point transform (layout: fb_core.cuh, FFTTables).
template <int M_>
inline void build_fft_tables_m(std::vector<cf>& twa, std::vector<cf>& twb) {
  typedef FFTGeom<M_> G;
  typedef FFTTables<M_> FT;
  auto W = [](long long e) {
    const double a = 2.0 * M_PI * (double)(((e % M_) + M_) % M_) / (double)M_;
    return mk((float)cos(a), (float)sin(a));
  };
  twa.assign(FT::TWA_WORDS > 0 ? FT::TWA_WORDS : 1, mk(1.f, 0.f));
  for (int j = 0; j < G::JA; j++) twa[(size_t)j * FT::TA] = G::Rb > 1 ? W((long long)(j / G::Rc) * G::Rc) : W((long long)j);
  twb.assign(FT::TWB_WORDS > 0 ? FT::TWB_WORDS : 1, mk(1.f, 0.f));
  if (G::Rb > 1)
    for (int gl = 0; gl < G::L; gl++) {
      const int nc = gl % G::Rc;
      twb[(size_t)gl * FT::TB + 0] = W((long long)nc * (gl / G::Rc));
      twb[(size_t)gl * FT::TB + 1] = W((long long)nc * (G::L / G::Rc));
      twb[(size_t)gl * FT::TB + 2] = W((long long)nc * G::Ra);
    }
}
inline bool build_fft_tables(int M, std::vector<cf>& twa, std::vector<cf>& twb) {
  switch (M) {
    case 64: build_fft_tables_m<64>(twa, twb); return true;
    case 128: build_fft_tables_m<128>(twa, twb); return true;
    case 256: build_fft_tables_m<256>(twa, twb); return true;
    case 512: build_fft_tables_m<512>(twa, twb); return true;
    case 1024: build_fft_tables_m<1024>(twa, twb); return true;
  }
  return false;
}

// Blocking matrix of a distortionless beamformer (beamformer/beamformer.cc:398-479 _calcBlockingMatrix, NC = 1):
// Gram-Schmidt on the first C - NC columns of  P = I - conj(v) v^T / ||v||^2  (zgeru, unconjugated rank-1 update), each
// column projected against the earlier block columns with zdotc (first argument conjugated) and normalised.
// Bm: [C][C - NC] row major.  false when C <= NC (the reference prints and fails).
inline bool blocking_matrix(const zd* v, int C, int NC, std::vector<zd>& Bm) {
  const int bs = C - NC;
  if (bs <= 0) return false;
  Bm.assign((size_t)C * bs, zd(0, 0));
  double nv = 0;
  for (int i = 0; i < C; i++) nv += std::norm(v[i]);
  std::vector<zd> vec(C);
  for (int id = 0; id < bs; id++) {
    for (int i = 0; i < C; i++) vec[i] = (i == id ? zd(1, 0) : zd(0, 0)) - std::conj(v[i]) * v[id] / nv;
    for (int jd = 0; jd < id; jd++) {
      zd ip(0, 0);
      for (int i = 0; i < C; i++) ip += std::conj(Bm[(size_t)i * bs + jd]) * vec[i];
      for (int i = 0; i < C; i++) vec[i] -= ip * Bm[(size_t)i * bs + jd];
    }
    double nn = 0;
    for (int i = 0; i < C; i++) nn += std::norm(vec[i]);
    nn = sqrt(nn);
    for (int i = 0; i < C; i++) Bm[(size_t)i * bs + id] = vec[i] / nn;
  }
  return true;
}

// Residue-major analysis taps: th[rho*TS + t] = h[rho + D t], t in [0, m R)  (chain_tile.cuh, polyphase_pair).
inline void build_analysis_taps(const double* h, int M, int m, int R, std::vector<float>& th) {
  const int D = M / R, mR = m * R, TS = tap_stride(mR);
  th.assign((size_t)D * TS, 0.f);
  if (!h) return;
  for (int rho = 0; rho < D; rho++)
    for (int t = 0; t < mR; t++) th[(size_t)rho * TS + t] = (float)h[rho + D * t];
}

// gp[k][q] = g[M-1-q + M k]  (polyphase(_M - m - 1, k), modulated.cc:649)
inline void build_synthesis_taps(const double* g, int M, int m, std::vector<float>& gp) {
  gp.resize((size_t)M * m);
  for (int k = 0; k < m; k++)
    for (int q = 0; q < M; q++) gp[(size_t)k * M + q] = (float)g[M - 1 - q + M * k];
}

// beamformerWeights::calcMainlobe, halfBandShift == false (beamformer/beamformer.cc:531-594).
// w: [B][C].
inline void ds_weights(const double* delays, double fs, int M, int C, std::vector<zd>& w) {
  const int B = M / 2 + 1;
  w.assign((size_t)B * C, zd(0, 0));
  for (int c = 0; c < C; c++) w[c] = zd(1.0 / C, 0.0);
  for (int s = 1; s < M / 2; s++)
    for (int c = 0; c < C; c++) {
      const double val = -2.0 * M_PI * s * delays[c] * fs / M;
      w[(size_t)s * C + c] = zd(cos(val) / C, sin(val) / C);
    }
  for (int c = 0; c < C; c++) {
    const double val = -M_PI * fs * delays[c];
    w[(size_t)(M / 2) * C + c] = zd(cos(val) / C, sin(val) / C);
  }
}

// Hermitian-extended conjugate weight table for the fused chain (see chain_tile.cuh header):
//   gam[c][k] = conj(w[k][c]) for 0 < k < M/2 ; gam[c][M-k] = w[k][c] ; real part only at k = 0, M/2,
// stored in the register order of the transform: element ((c*(V/2) + r2)*L + gl)*2 + i holds bin
// index_of_spec(gl, 2 r2 + i), so that a lane's two consecutive registers are one 16-byte load and the lanes of a
// group read consecutive 16-byte words.
template <int M_>
inline void build_chain_weight_table_m(const zd* w, int C, int Cpad, std::vector<cf>& gam) {
  typedef FFTGeom<M_> G;
  const int M = M_;
  gam.assign((size_t)Cpad * M, mk(0.f, 0.f));
  std::vector<cf> row(M);
  for (int c = 0; c < C; c++) {
    row[0] = mk((float)w[c].real(), 0.f);
    row[M / 2] = mk((float)w[(size_t)(M / 2) * C + c].real(), 0.f);
    for (int k = 1; k < M / 2; k++) {
      const zd v = w[(size_t)k * C + c];
      row[k] = mk((float)v.real(), (float)-v.imag());
      row[M - k] = mk((float)v.real(), (float)v.imag());
    }
    for (int gl = 0; gl < G::L; gl++)
      for (int r = 0; r < G::V; r++)
        gam[(((size_t)c * (G::V / 2) + r / 2) * G::L + gl) * 2 + (r & 1)] = row[G::index_of_spec(gl, r)];
  }
}
// bin index of every slot of one channel's row of the chain weight table (for the device-side table builder)
template <int M_>
inline void build_weight_binmap_m(std::vector<int>& map) {
  typedef FFTGeom<M_> G;
  map.assign(M_, 0);
  for (int gl = 0; gl < G::L; gl++)
    for (int r = 0; r < G::V; r++) map[((size_t)(r / 2) * G::L + gl) * 2 + (r & 1)] = G::index_of_spec(gl, r);
}
inline bool build_weight_binmap(int M, std::vector<int>& map) {
  switch (M) {
    case 64: build_weight_binmap_m<64>(map); return true;
    case 128: build_weight_binmap_m<128>(map); return true;
    case 256: build_weight_binmap_m<256>(map); return true;
    case 512: build_weight_binmap_m<512>(map); return true;
    case 1024: build_weight_binmap_m<1024>(map); return true;
  }
  return false;
}
inline bool build_chain_weight_table(const zd* w, int M, int C, int Cpad, std::vector<cf>& gam) {
  switch (M) {
    case 64: build_chain_weight_table_m<64>(w, C, Cpad, gam); return true;
    case 128: build_chain_weight_table_m<128>(w, C, Cpad, gam); return true;
    case 256: build_chain_weight_table_m<256>(w, C, Cpad, gam); return true;
    case 512: build_chain_weight_table_m<512>(w, C, Cpad, gam); return true;
    case 1024: build_chain_weight_table_m<1024>(w, C, Cpad, gam); return true;
  }
  return false;
}

// SubbandMVDR::setDiffuseNoiseModel (beamformer/beamformer.cc:2486-2553): Gamma_mn = sinc(2 fs s d_mn/(M c)),
// normalised sinc, unit diagonal.  Rn: [B][C][C].
inline double norm_sinc(double x) {
  if (fabs(x) < 1e-8) return 1.0 - (M_PI * M_PI * x * x) / 6.0;
  return sin(M_PI * x) / (M_PI * x);
}
inline void diffuse_model(const double* micpos, int C, double fs, double sspeed, int M, std::vector<zd>& Rn) {
  const int B = M / 2 + 1;
  Rn.assign((size_t)B * C * C, zd(0, 0));
  for (int s = 0; s < B; s++) {
    const double omega_d_c = 2.0 * fs * s / (M * sspeed);
    for (int a = 0; a < C; a++)
      for (int b = 0; b < C; b++) {
        double v = 1.0;
        if (a != b) {
          const double dx = micpos[a * 3] - micpos[b * 3], dy = micpos[a * 3 + 1] - micpos[b * 3 + 1],
                       dz = micpos[a * 3 + 2] - micpos[b * 3 + 2];
          v = norm_sinc(omega_d_c * sqrt(dx * dx + dy * dy + dz * dz));
        }
        Rn[((size_t)s * C + a) * C + b] = zd(v, 0.0);
      }
  }
}

// Split every recording's nblk output frames into chunks of at most `chunk` frames.
inline void build_work(const std::vector<RecDesc>& recs, int chunk, std::vector<WorkItem>& work) {
  work.clear();
  for (size_t r = 0; r < recs.size(); r++)
    for (int j0 = 0; j0 < recs[r].nblk; j0 += chunk) {
      WorkItem w;
      w.rec = (int)r; w.j0 = j0; w.nj = recs[r].nblk - j0 < chunk ? recs[r].nblk - j0 : chunk;
      work.push_back(w);
    }
}

}  // namespace btk
