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

// calcNullBeamformer (beamformer/beamformer.cc:315-397): constraint matrix Cm = [wt | pWj_0 | ...] ([C][NC]),
// wt <- Cm (Cm^H Cm)^{-1} e_0.  NC == 2 goes through putInverseMat22 (:202-242: plain 2 x 2 inverse, 0.01 added to the
// diagonal when |det| < 1e-7); larger NC through pseudoinverse() (:253-305, single-precision SVD) in the reference --
// here a double-precision Gauss-Jordan inverse, which agrees with it to the float SVD's own rounding.
inline bool null_beamformer(zd* wt, const zd* pWj /*[NC-1][C]*/, int C, int NC) {
  std::vector<zd> Cm((size_t)C * NC), A((size_t)NC * NC, zd(0, 0));
  for (int i = 0; i < C; i++) {
    Cm[(size_t)i * NC] = wt[i];
    for (int j = 1; j < NC; j++) Cm[(size_t)i * NC + j] = pWj[(size_t)(j - 1) * C + i];
  }
  for (int a = 0; a < NC; a++)
    for (int b = 0; b < NC; b++) {
      zd s(0, 0);
      for (int i = 0; i < C; i++) s += std::conj(Cm[(size_t)i * NC + a]) * Cm[(size_t)i * NC + b];
      A[(size_t)a * NC + b] = s;
    }
  std::vector<zd> v(NC);                      // first column of A^{-1}
  if (NC == 2) {
    zd m00 = A[0], m01 = A[1], m10 = A[2], m11 = A[3];
    zd det = m00 * m11 - m01 * m10;
    if (std::abs(det) < 1.0e-7) { m00 += 0.01; m11 += 0.01; det = m00 * m11 - m01 * m10; }
    v[0] = m11 / det;
    v[1] = -(m10 / det);
  } else {
    std::vector<zd> aug((size_t)NC * (NC + 1));
    for (int a = 0; a < NC; a++) {
      for (int b = 0; b < NC; b++) aug[(size_t)a * (NC + 1) + b] = A[(size_t)a * NC + b];
      aug[(size_t)a * (NC + 1) + NC] = a == 0 ? zd(1, 0) : zd(0, 0);
    }
    for (int col = 0; col < NC; col++) {
      int piv = col;
      for (int a = col + 1; a < NC; a++)
        if (std::abs(aug[(size_t)a * (NC + 1) + col]) > std::abs(aug[(size_t)piv * (NC + 1) + col])) piv = a;
      if (std::abs(aug[(size_t)piv * (NC + 1) + col]) == 0.0) return false;
      if (piv != col)
        for (int b = 0; b <= NC; b++) std::swap(aug[(size_t)piv * (NC + 1) + b], aug[(size_t)col * (NC + 1) + b]);
      const zd ip = zd(1, 0) / aug[(size_t)col * (NC + 1) + col];
      for (int b = 0; b <= NC; b++) aug[(size_t)col * (NC + 1) + b] *= ip;
      for (int a = 0; a < NC; a++) {
        if (a == col) continue;
        const zd f = aug[(size_t)a * (NC + 1) + col];
        for (int b = 0; b <= NC; b++) aug[(size_t)a * (NC + 1) + b] -= f * aug[(size_t)col * (NC + 1) + b];
      }
    }
    for (int a = 0; a < NC; a++) v[a] = aug[(size_t)a * (NC + 1) + NC];
  }
  for (int i = 0; i < C; i++) {
    zd s(0, 0);
    for (int j = 0; j < NC; j++) s += Cm[(size_t)i * NC + j] * v[j];
    wt[i] = s;
  }
  return true;
}

// beamformerWeights::calcMainlobeN, halfBandShift == false (beamformer/beamformer.cc:632-735; calcMainlobe2 :603-623 is
// NC == 2): distortionless towards delaysT with nulls towards the NC - 1 interferers delaysJ [NC-1][C].
// ta <- the delay-and-sum weights of the target (what the reference keeps as the array manifold _ta), w <- the
// quiescent weights.  Restated statement for statement, including the reference's handling of bin M/2 (:722-734): inside
// the channel loop element c is first rescaled, then overwritten with the LAST interferer's steering value / C, and the
// null beamformer is re-run on the whole vector after every channel with the interferer vectors left over from bin
// M/2 - 1.  Only bins 0..M/2 are produced (SubbandDS::next mirrors the rest, :1189-1194).
inline bool null_weights(const double* delaysT, const double* delaysJ, double fs, int M, int C, int NC, std::vector<zd>& ta,
                         std::vector<zd>& w) {
  ds_weights(delaysT, fs, M, C, ta);
  w = ta;
  std::vector<zd> pWj((size_t)(NC - 1) * C, zd(0, 0));
  for (int s = 1; s < M / 2; s++) {
    zd* vec = &w[(size_t)s * C];
    for (int c = 0; c < C; c++) {
      vec[c] *= (double)C;
      for (int n = 0; n < NC - 1; n++) {
        const double valJ = -2.0 * M_PI * s * fs * delaysJ[(size_t)n * C + c] / M;
        pWj[(size_t)n * C + c] = zd(cos(valJ), sin(valJ));
      }
    }
    if (!null_beamformer(vec, pWj.data(), C, NC)) return false;
  }
  zd* vec = &w[(size_t)(M / 2) * C];
  for (int c = 0; c < C; c++) {
    vec[c] *= (double)C;
    for (int n = 0; n < NC - 1; n++) {
      const double val = -M_PI * fs * delaysJ[(size_t)n * C + c];
      vec[c] = zd(cos(val), sin(val)) / (double)C;
    }
    if (!null_beamformer(vec, pWj.data(), C, NC)) return false;
  }
  return true;
}

// Far-field delays of the shipped C++ driver (src/superdirectiveBeamformer.cc:118-137 calcDelaysPolar2): direction
// cosines, coordinates and the quotient are computed in SINGLE precision there (float locals, SOUNDSPEED 343740.0 mm/s),
// then widened into the gsl_vector; restated with the same types so the delays agree bit for bit.
inline void delays_polar2(float azimuth, float elevation, const double* micpos /*[n][3] mm*/, int n, double* delays) {
  const float c_x = -sinf(elevation) * cosf(azimuth);
  const float c_y = -sinf(elevation) * sinf(azimuth);
  const float c_z = -cosf(elevation);
  for (int i = 0; i < n; i++) {
    const float x = (float)micpos[3 * i], y = (float)micpos[3 * i + 1], z = (float)micpos[3 * i + 2];
    const float t = (float)((double)(c_x * x + c_y * y + c_z * z) / 343740.0);
    delays[i] = t;
  }
}
// calcAllDelays (beamformer/beamformer.cc:1214-1231): distance of every microphone from the ORIGIN over the speed of
// sound, minus the value of the middle element -- the source position (x, y, z) is accepted and ignored, as in the
// reference (:1219-1223 never read it).
inline void all_delays(double /*x*/, double /*y*/, double /*z*/, const double* micpos, int n, double* delays) {
  for (int c = 0; c < n; c++) {
    const double xm = micpos[3 * c], ym = micpos[3 * c + 1], zm = micpos[3 * c + 2];
    delays[c] = sqrt(xm * xm + ym * ym + zm * zm) / 343740.0;
  }
  const double mid = delays[n / 2];
  for (int c = 0; c < n; c++) delays[c] -= mid;
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

// Persistent schedule of the warp-specialised chain (chain_ws.cuh::WsSegs): first item of each of `ncta` CTAs such that no
// CTA runs more than B iterations, B minimal.  Recording r has prefix[r+1] - prefix[r] items of q frames (the last one
// may be shorter); a segment of nj frames costs ceil((nj + H) / W) iterations; a CTA walks on into the next recording
// while its budget lasts.  Returns B; `begin` gets ncta + 1 item indices (trailing CTAs may be empty).
inline int balance_ctas(const std::vector<RecDesc>& recs, const std::vector<int>& prefix, int q, int W, int H, int ncta,
                        std::vector<int>& begin) {
  const int n = (int)recs.size();
  auto walk = [&](long long B, std::vector<int>* out) -> int {
    int used = 0, r = 0, j = 0;                      // next recording / next frame inside it
    while (r < n && recs[r].nblk == 0) r++;
    if (out) out->assign(1, r < n ? prefix[r] : prefix[n]);
    while (r < n) {
      long long b = B;
      for (;;) {
        // frames this CTA can still take from recording r: whole items, b W - H frames at most
        const long long room = b * W - H;
        if (room < 1) break;
        const int rem = recs[r].nblk - j;
        long long take = rem <= room ? rem : room / q * q;
        if (take < 1) break;
        b -= (take + H + W - 1) / W;
        j += (int)take;
        if (j >= recs[r].nblk) {
          r++; j = 0;
          while (r < n && recs[r].nblk == 0) r++;
          if (r >= n) break;
        } else {
          break;                                      // budget exhausted inside the recording
        }
      }
      used++;
      if (out) out->push_back(r < n ? prefix[r] + j / q : prefix[n]);
      if (used > 2 * ncta + 8) break;                // budget too small to ever finish
    }
    return used;
  };
  long long total = 0;
  for (int r = 0; r < n; r++) total += recs[r].nblk;
  long long lo = 1, hi = (total + H + W - 1) / W + n + 1;      // hi: one CTA takes everything
  while (walk(hi, nullptr) > ncta) hi *= 2;
  while (lo < hi) {
    const long long mid = (lo + hi) / 2;
    if (walk(mid, nullptr) <= ncta) hi = mid; else lo = mid + 1;
  }
  walk(lo, &begin);
  begin.resize((size_t)ncta + 1, prefix[n]);
  return (int)lo;
}

// Largest number of iterations any CTA runs when the items of the launch are split evenly (what WsSegs does without
// p.cta_begin): the yardstick for balance_ctas.
inline long long even_split_iterations(const std::vector<RecDesc>& recs, const std::vector<int>& prefix, int q, int W, int H, int ncta) {
  const int n = (int)recs.size();
  const long long total = prefix[n];
  long long worst = 0;
  int r = 0;
  for (int c = 0; c < ncta; c++) {
    long long i = total * c / ncta;
    const long long i1 = total * (c + 1) / ncta;
    long long its = 0;
    while (i < i1) {
      while (prefix[r + 1] <= i) r++;
      const long long e = prefix[r + 1] < i1 ? prefix[r + 1] : i1;
      const long long j0 = (i - prefix[r]) * q, j1 = (e - prefix[r]) * q;
      const long long nj = (j1 < recs[r].nblk ? j1 : recs[r].nblk) - j0;
      its += (nj + H + W - 1) / W;
      i = e;
    }
    if (its > worst) worst = its;
  }
  return worst;
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
