// fb_core.cuh -- register/shared-memory FFT building blocks of the oversampled DFT filter bank.
//
// Everything here is __host__ __device__ so the exact same code runs (a) inside the sm_100a
// kernels and (b) under the sequential host emulator used by the CPU-only tests
// (tests/emu/): there is no GPU in the build container, so index arithmetic is validated on
// the host before any GPU time is spent.  The emulator is a test harness for THIS code, not a
// fallback: the product library only ever launches the CUDA kernels.
//
// Math restated (reference btk/modulated/modulated.cc:412-452, 595-664):
//   analysis  X_i[s] = sum_q u_i[q] e^{+j 2 pi s q / M}   ("backward", SIGN=+1)
//   synthesis v[q]   = Re sum_s Y[s] e^{-j 2 pi s q / M}  ("forward",  SIGN=-1)
// Two real sequences are packed into one complex transform (frames i and i+1 of one channel
// on the analysis side; two Hermitian spectra on the synthesis side), see DESIGN.md.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BTK_HD __host__ __device__ __forceinline__
#define BTK_UNROLL _Pragma("unroll")
#define BTK_UNROLL_N(n) _Pragma(BTK_STR(unroll n))
#define BTK_STR(x) #x
#if defined(__CUDA_ARCH__)
#define BTK_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#else
#define BTK_PREFETCH_L2(ptr) ((void)(ptr))
#endif
#else
#define BTK_PREFETCH_L2(ptr) ((void)(ptr))
#define BTK_HD inline
#define BTK_UNROLL
#define BTK_UNROLL_N(n)
#ifndef BTK_HOST_VECTOR_TYPES
#define BTK_HOST_VECTOR_TYPES
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
#endif
static inline float fmaf_host(float a, float b, float c) { return fmaf(a, b, c); }
#endif

namespace btk {

typedef float2 cf;

BTK_HD cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }

// ---------------------------------------------------------------------------------------------
// Complex arithmetic on the Blackwell packed-FP32 pipe.  sm_100 executes add/mul/fma.f32x2 on an aligned
// register pair (FADD2 / FMUL2 / FFMA2) and its operand modifiers swap the halves (.LO_HI), negate either
// half and broadcast a scalar register to both halves (.F32), so a complex add is ONE instruction and a
// complex multiply(-accumulate) is TWO -- ptxas folds the swaps / negations / broadcasts written below as
// plain C into those modifiers (checked with cuobjdump).  Measured on B200 (tools/ubench):
// 128 flop-lanes/clk/SM for the packed forms against 64-107 for the scalar forms, at half the issue slots.
// The host versions (sequential emulator of the CPU tests) are the same formulas in scalar arithmetic.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define BTK_U64(x) (*reinterpret_cast<unsigned long long*>(&(x)))
__device__ __forceinline__ cf pk_add(cf a, cf b) { cf r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(BTK_U64(r)) : "l"(BTK_U64(a)), "l"(BTK_U64(b))); return r; }
__device__ __forceinline__ cf pk_sub(cf a, cf b) { cf r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(BTK_U64(r)) : "l"(BTK_U64(a)), "l"(BTK_U64(b))); return r; }
__device__ __forceinline__ cf pk_mul(cf a, cf b) { cf r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(BTK_U64(r)) : "l"(BTK_U64(a)), "l"(BTK_U64(b))); return r; }
__device__ __forceinline__ cf pk_fma(cf a, cf b, cf c) { cf r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(BTK_U64(r)) : "l"(BTK_U64(a)), "l"(BTK_U64(b)), "l"(BTK_U64(c))); return r; }
#else
BTK_HD cf pk_add(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
BTK_HD cf pk_sub(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
BTK_HD cf pk_mul(cf a, cf b) { return mk(a.x * b.x, a.y * b.y); }
BTK_HD cf pk_fma(cf a, cf b, cf c) { return mk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
BTK_HD cf cswap(cf a) { return mk(a.y, a.x); }
BTK_HD cf cadd(cf a, cf b) { return pk_add(a, b); }
BTK_HD cf csub(cf a, cf b) { return pk_sub(a, b); }
// a * b = a (b.x, b.x) + swap(a) (-b.y, b.y)
BTK_HD cf cmul(cf a, cf b) { return pk_fma(cswap(a), mk(-b.y, b.y), pk_mul(a, mk(b.x, b.x))); }
// a * conj(b) = a (b.x, b.x) + swap(a) (b.y, -b.y)
BTK_HD cf cmulc(cf a, cf b) { return pk_fma(cswap(a), mk(b.y, -b.y), pk_mul(a, mk(b.x, b.x))); }
// acc += a * b
BTK_HD void cfma(cf& acc, cf a, cf b) {
  acc = pk_fma(a, mk(b.x, b.x), acc);
  acc = pk_fma(cswap(a), mk(-b.y, b.y), acc);
}
// acc += a * s  (real scalar s on both halves)
BTK_HD cf cfma_real(cf a, float s, cf acc) { return pk_fma(a, mk(s, s), acc); }
// multiply by (j*S)
template <int S> BTK_HD cf mulj(cf a) { return S > 0 ? mk(-a.y, a.x) : mk(a.y, -a.x); }
// multiply by (c + j*S*s)
template <int S> BTK_HD cf mulw(cf a, float c, float s) {
  return S > 0 ? pk_fma(cswap(a), mk(-s, s), pk_mul(a, mk(c, c))) : pk_fma(cswap(a), mk(s, -s), pk_mul(a, mk(c, c)));
}
// multiply by table twiddle w = e^{+j theta}; for S<0 use conj(w)
template <int S> BTK_HD cf multw(cf a, cf w) { return S > 0 ? cmul(a, w) : cmulc(a, w); }

#define BTK_SQRT1_2 0.70710678118654752440f
#define BTK_COS_PI_8 0.92387953251128675613f
#define BTK_SIN_PI_8 0.38268343236508977173f
#define BTK_COS_PI_16 0.98078528040323044913f
#define BTK_SIN_PI_16 0.19509032201612826785f

// ---------------------------------------------------------------------------------------------
// Small in-register DFTs, natural-order in and out:  X[k] = sum_n x[n] e^{S j 2 pi n k / R}.
// ---------------------------------------------------------------------------------------------
template <int S> BTK_HD void dft2(cf& a, cf& b) { cf t = csub(a, b); a = cadd(a, b); b = t; }

template <int S> BTK_HD void dft4(cf& x0, cf& x1, cf& x2, cf& x3) {
  cf t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mulj<S>(csub(x1, x3));
  x0 = cadd(t0, t2); x1 = cadd(t1, t3); x2 = csub(t0, t2); x3 = csub(t1, t3);
}

template <int S> BTK_HD void dft8(cf* v) {
  // decimation in time: even/odd 4-point DFTs, then W8^{S k}
  cf e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  cf o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4<S>(e0, e1, e2, e3);
  dft4<S>(o0, o1, o2, o3);
  // W8^1 = (1 + jS)/sqrt2 ; W8^2 = jS ; W8^3 = (-1 + jS)/sqrt2
  cf t1 = mulw<S>(o1, BTK_SQRT1_2, BTK_SQRT1_2);
  cf t2 = mulj<S>(o2);
  cf t3 = mulw<S>(o3, -BTK_SQRT1_2, BTK_SQRT1_2);
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
  v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
  v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

template <int S> BTK_HD void dft16(cf* v) {
  // n = 4 n1 + n2, k = k1 + 4 k2
  cf a[4][4];  // a[n2][k1]
  BTK_UNROLL
  for (int n2 = 0; n2 < 4; n2++) {
    cf x0 = v[n2], x1 = v[4 + n2], x2 = v[8 + n2], x3 = v[12 + n2];
    dft4<S>(x0, x1, x2, x3);
    a[n2][0] = x0; a[n2][1] = x1; a[n2][2] = x2; a[n2][3] = x3;
  }
  // twiddles W16^{S n2 k1}
  a[1][1] = mulw<S>(a[1][1], BTK_COS_PI_8, BTK_SIN_PI_8);                    // W^1
  a[1][2] = mulw<S>(a[1][2], BTK_SQRT1_2, BTK_SQRT1_2);                      // W^2
  a[1][3] = mulw<S>(a[1][3], BTK_SIN_PI_8, BTK_COS_PI_8);                    // W^3
  a[2][1] = mulw<S>(a[2][1], BTK_SQRT1_2, BTK_SQRT1_2);                      // W^2
  a[2][2] = mulj<S>(a[2][2]);                                               // W^4
  a[2][3] = mulw<S>(a[2][3], -BTK_SQRT1_2, BTK_SQRT1_2);                     // W^6
  a[3][1] = mulw<S>(a[3][1], BTK_SIN_PI_8, BTK_COS_PI_8);                    // W^3
  a[3][2] = mulw<S>(a[3][2], -BTK_SQRT1_2, BTK_SQRT1_2);                     // W^6
  a[3][3] = mulw<S>(a[3][3], -BTK_COS_PI_8, -BTK_SIN_PI_8);                  // W^9
  BTK_UNROLL
  for (int k1 = 0; k1 < 4; k1++) {
    cf x0 = a[0][k1], x1 = a[1][k1], x2 = a[2][k1], x3 = a[3][k1];
    dft4<S>(x0, x1, x2, x3);
    v[k1] = x0; v[k1 + 4] = x1; v[k1 + 8] = x2; v[k1 + 12] = x3;
  }
}

// 32-point DFT: even / odd 16-point DFTs, then W32^{S k}
template <int S> BTK_HD void dft32(cf* v) {
  cf e[16], o[16];
  BTK_UNROLL
  for (int i = 0; i < 16; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
  dft16<S>(e);
  dft16<S>(o);
  // cos / sin of 2 pi k / 32, k = 0..15
  const float c32[16] = {1.f, 0.98078528040323044913f, BTK_COS_PI_8, 0.83146961230254523708f, BTK_SQRT1_2, 0.55557023301960222474f,
                         BTK_SIN_PI_8, 0.19509032201612826785f, 0.f, -0.19509032201612826785f, -BTK_SIN_PI_8,
                         -0.55557023301960222474f, -BTK_SQRT1_2, -0.83146961230254523708f, -BTK_COS_PI_8, -0.98078528040323044913f};
  const float s32[16] = {0.f, 0.19509032201612826785f, BTK_SIN_PI_8, 0.55557023301960222474f, BTK_SQRT1_2, 0.83146961230254523708f,
                         BTK_COS_PI_8, 0.98078528040323044913f, 1.f, 0.98078528040323044913f, BTK_COS_PI_8,
                         0.83146961230254523708f, BTK_SQRT1_2, 0.55557023301960222474f, BTK_SIN_PI_8, 0.19509032201612826785f};
  BTK_UNROLL
  for (int k = 0; k < 16; k++) {
    const cf t = k == 0 ? o[0] : (k == 8 ? mulj<S>(o[8]) : mulw<S>(o[k], c32[k], s32[k]));
    v[k] = cadd(e[k], t);
    v[k + 16] = csub(e[k], t);
  }
}

template <int R, int S> struct Dft;
template <int S> struct Dft<1, S> { static BTK_HD void run(cf*) {} };
template <int S> struct Dft<2, S> { static BTK_HD void run(cf* v) { dft2<S>(v[0], v[1]); } };
template <int S> struct Dft<4, S> { static BTK_HD void run(cf* v) { dft4<S>(v[0], v[1], v[2], v[3]); } };
template <int S> struct Dft<8, S> { static BTK_HD void run(cf* v) { dft8<S>(v); } };
template <int S> struct Dft<16, S> { static BTK_HD void run(cf* v) { dft16<S>(v); } };
template <int S> struct Dft<32, S> { static BTK_HD void run(cf* v) { dft32<S>(v); } };

// ---------------------------------------------------------------------------------------------
// M-point transform shared by the L lanes of a group, V = M/L values per lane.
//   M = Ra * Rb * Rc  with Rc == Ra (so the output register layout equals the input layout):
//   n = na*(Rb*Rc) + nb*Rc + nc ,  k = ka + Ra*kb + Ra*Rb*kc
//   pass A: radix-Ra over na   (owner: j  = nb*Rc + nc  = gl + L*rep)   * W_{Ra Rb}^{nb ka}
//   pass B: radix-Rb over nb   (owner: iB = ka*Rc + nc  = gl + L*rep)   * W_M^{nc (ka + Ra kb)}
//   pass C: radix-Rc over nc   (owner: iC = ka + Ra*kb  = gl + L*rep)
// Register layout (input and output): value r = rep*Ra + e  <->  index  (gl + L*rep) + (M/Ra)*e.
// Passes exchange data through a per-group shared-memory buffer of XBUF complex words.
// ---------------------------------------------------------------------------------------------
template <int M_> struct FFTPlan;
template <> struct FFTPlan<64>   { static constexpr int M = 64,   Ra = 8,  Rb = 1, Rc = 8,  V = 8,  L = 8;  };
#ifdef BTK_FFT128_1EXCH    // A/B: 16 x 8 with one exchange measured 5 % SLOWER than the three-pass plan at M = 128 (0.293 vs 0.278 ms)
template <> struct FFTPlan<128>  { static constexpr int M = 128,  Ra = 16, Rb = 1, Rc = 8,  V = 16, L = 8;  };
#else
template <> struct FFTPlan<128>  { static constexpr int M = 128,  Ra = 8,  Rb = 2, Rc = 8,  V = 8,  L = 16; };
#endif
template <> struct FFTPlan<256>  { static constexpr int M = 256,  Ra = 16, Rb = 1, Rc = 16, V = 16, L = 16; };
#ifdef BTK_FFT512_3PASS    // A/B: the three-pass plan of the first sessions (two exchanges, 32 lanes per transform)
template <> struct FFTPlan<512>  { static constexpr int M = 512,  Ra = 16, Rb = 2, Rc = 16, V = 16, L = 32; };
#else
// 512 = 32 x 16 with ONE exchange: radix-32 in registers, then radix-16; 16 lanes per transform (two channels per warp).
// Ra != Rc: the spectrum comes out in a register order different from the time-domain one (see FFTGeom::index_of_spec),
// and the synthesis direction runs the transposed flow (radix-16 first, GroupFFT::inv_*).
template <> struct FFTPlan<512>  { static constexpr int M = 512,  Ra = 32, Rb = 1, Rc = 16, V = 32, L = 16; };
#endif
#ifdef BTK_FFT1024_3PASS   // A/B: the three-pass plan of the first sessions (two exchanges)
template <> struct FFTPlan<1024> { static constexpr int M = 1024, Ra = 16, Rb = 4, Rc = 16, V = 32, L = 32; };
#else
template <> struct FFTPlan<1024> { static constexpr int M = 1024, Ra = 32, Rb = 1, Rc = 32, V = 32, L = 32; };   // one exchange
#endif

template <int M_> struct FFTGeom {
  typedef FFTPlan<M_> P;
  static constexpr int M = P::M, Ra = P::Ra, Rb = P::Rb, Rc = P::Rc, V = P::V, L = P::L;
  static constexpr int NG = 32 / L;            // groups per warp
  static constexpr int JA = M / Ra;            // owners of pass A per transform (= Rb*Rc)
  static constexpr int RepA = V / Ra;          // radix-Ra DFTs per lane in pass A (and C)
  static constexpr int RepC = V / Rc;          // radix-Rc DFTs per lane in pass C
  static constexpr int RepB = (Rb > 1) ? V / Rb : 0;
  static constexpr bool ASYM = (Rb == 1) && (Ra != Rc);   // time-domain and spectrum register orders differ
  static constexpr int PadA = (Rc % 32 == 0) ? 0 : Rc;  // S1 == Rc (mod 32) keeps pass-B reads conflict free
  static constexpr int S1 = JA + PadA;         // row stride (complex words) of exchange 1: idx = ka*S1 + j
  static constexpr int S2 = Rc + 1;            // row stride of exchange 2: idx = iC*S2 + nc
  static constexpr int X1 = (Rb > 1) ? Ra * S1 : 0;
  static constexpr int X2 = (M / Rc) * S2;
  static constexpr int XBUF = X1 > X2 ? X1 : X2;   // complex words per group
  // register <-> transform index
  // time-domain order (input of the analysis transform, output of the synthesis transform)
  static BTK_HD int index_of(int gl, int r) { return (gl + L * (r / Ra)) + JA * (r % Ra); }
  // spectrum order (output of the analysis transform, input of the synthesis transform): value r = rep*Rc + kc of a
  // lane is bin (gl + L rep) + Ra kc.  Identical to index_of unless ASYM.
  static BTK_HD int index_of_spec(int gl, int r) { return ASYM ? (gl + L * (r / Rc)) + Ra * (r % Rc) : index_of(gl, r); }
};

// Twiddle seeds (built on the host by host_tables.h::build_fft_tables, copied to shared memory by every tile
// program).  A lane needs Ra-1 pass-A twiddles and, for the three-pass sizes, V pass-B twiddles per transform; all
// of them are powers of a few lane constants, so only those seeds are read and the powers are formed in registers
// with packed complex multiplies (the kernels are bound by shared-memory bandwidth, not by FP32 issue):
//   pass A, owner j = gl + L rep in [0, JA):   twa[j] = W_M^j (two-pass) or W_M^{(j / Rc) Rc} (three-pass);
//                                              twiddle of output ka = twa[j]^ka
//   pass B, lane gl (nc = gl % Rc):            twb[4 gl + 0] = W_M^{nc (gl / Rc)},  [1] = W_M^{nc L / Rc},  [2] = W_M^{nc Ra}
//                                              twiddle of (rep, kb) = [0] * [1]^rep * [2]^kb  =  W_M^{nc (ka + Ra kb)},
//                                              ka = (gl + L rep) / Rc
// with W_M = e^{+j 2 pi / M}.
template <int M_> struct FFTTables {
  typedef FFTGeom<M_> G;
  static constexpr int TA = 1;
  static constexpr int TWA_WORDS = G::JA * TA;
  static constexpr int TB = 4;
  static constexpr int TWB_WORDS = G::Rb > 1 ? G::L * TB : 0;
  static_assert(G::Rb == 1 || G::Rb == 2 || G::Rb == 4, "pass B radix");
  static_assert(G::Rb == 1 || G::L % G::Rc == 0, "pass-B owners must keep nc per lane");
};

// One lane's share of the pass structure.
// The three steps must be separated by a group-wide barrier (__syncwarp on the device).
template <int M_, int S> struct GroupFFT {
  typedef FFTGeom<M_> G;
  typedef FFTTables<M_> FT;

  // step 1: pass A (+ twiddle) and scatter into the exchange buffer, for PP independent transforms of the same
  // lane (registers v + pp*V, exchange buffers xb + pp*XBUF) that share one set of twiddles.
  template <int PP>
  static BTK_HD void step1_multi(cf* v, int gl, cf* xb, const cf* twa) {
    BTK_UNROLL
    for (int rep = 0; rep < G::RepA; rep++) {
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) Dft<G::Ra, S>::run(v + pp * G::V + rep * G::Ra);
      const int j = gl + G::L * rep;
      // W^{ka} from the seed W = twa[j] by a product tree of depth <= 4.  The packed multiplies are cheaper than the
      // shared-memory bandwidth a table read would take.
      // (radix 32: only W^1..W^3 and W^4, W^8, .., W^28 are kept -- 11 values instead of 31 -- and W^ka = W^{ka & 3} W^{ka & ~3}
      // is formed when it is applied)
      constexpr int NLO = G::Ra > 16 ? 4 : G::Ra;               // w[1 .. NLO-1] directly
      cf w[NLO + 1];
      w[1] = twa[j * FT::TA];
      BTK_UNROLL
      for (int ka = 2; ka < NLO; ka++) {
        const int hi = ka >= 8 ? 8 : (ka >= 4 ? 4 : 2);        // largest power of two <= ka
        w[ka] = (ka == hi) ? cmul(w[ka / 2], w[ka / 2]) : cmul(w[hi], w[ka - hi]);
      }
      cf wh[G::Ra > 16 ? G::Ra / 4 : 1];                        // wh[i] = W^{4 i}
      if (G::Ra > 16) {
        wh[1] = cmul(w[2], w[2]);
        BTK_UNROLL
        for (int i = 2; i < G::Ra / 4; i++) {
          const int hi = i >= 4 ? 4 : 2;
          wh[i] = (i == hi) ? cmul(wh[i / 2], wh[i / 2]) : cmul(wh[hi], wh[i - hi]);
        }
      }
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        cf* p = v + pp * G::V + rep * G::Ra;
        cf* x = xb + pp * G::XBUF;
        BTK_UNROLL
        for (int ka = 1; ka < G::Ra; ka++) {
          if (G::Ra > 16) {
            const cf wk = (ka & 3) == 0 ? wh[ka >> 2] : ((ka >> 2) == 0 ? w[ka & 3] : cmul(w[ka & 3], wh[ka >> 2]));
            p[ka] = multw<S>(p[ka], wk);
          } else {
            p[ka] = multw<S>(p[ka], w[ka]);
          }
        }
        BTK_UNROLL
        for (int ka = 0; ka < G::Ra; ka++) x[ka * (G::Rb > 1 ? G::S1 : G::S2) + j] = p[ka];
      }
    }
  }
  static BTK_HD void step1(cf* v, int gl, cf* xb, const cf* twa) { step1_multi<1>(v, gl, xb, twa); }

  // The same pass A split in two for transforms that take turns on ONE exchange buffer (the warp-specialised chain,
  // chain_ws.cuh): step1_twiddle leaves the twiddled pass-A outputs of PP transforms in registers, step1_scatter writes
  // one of them to the buffer, step3_gather reads one back and step3_dft runs the final radix-Rc transforms.
  // Two-pass plans with the symmetric layout only (Rb == 1, Ra <= 16).
  template <int PP>
  static BTK_HD void step1_twiddle(cf* v, int gl, const cf* twa) {
    static_assert(G::Rb == 1 && G::Ra <= 16, "two-pass plans with a full product tree");
    BTK_UNROLL
    for (int rep = 0; rep < G::RepA; rep++) {
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) Dft<G::Ra, S>::run(v + pp * G::V + rep * G::Ra);
      const int j = gl + G::L * rep;
      cf w[G::Ra + 1];
      w[1] = twa[j * FT::TA];
      BTK_UNROLL
      for (int ka = 2; ka < G::Ra; ka++) {
        const int hi = ka >= 8 ? 8 : (ka >= 4 ? 4 : 2);
        w[ka] = (ka == hi) ? cmul(w[ka / 2], w[ka / 2]) : cmul(w[hi], w[ka - hi]);
      }
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        cf* p = v + pp * G::V + rep * G::Ra;
        BTK_UNROLL
        for (int ka = 1; ka < G::Ra; ka++) p[ka] = multw<S>(p[ka], w[ka]);
      }
    }
  }
  static BTK_HD void step1_scatter(const cf* v, int gl, cf* xb) {
    BTK_UNROLL
    for (int rep = 0; rep < G::RepA; rep++) {
      const int j = gl + G::L * rep;
      BTK_UNROLL
      for (int ka = 0; ka < G::Ra; ka++) xb[ka * G::S2 + j] = v[rep * G::Ra + ka];
    }
  }
  static BTK_HD void step3_gather(cf* v, int gl, const cf* xb) {
    BTK_UNROLL
    for (int rep = 0; rep < G::RepC; rep++) {
      const int iC = gl + G::L * rep;
      BTK_UNROLL
      for (int nc = 0; nc < G::Rc; nc++) v[rep * G::Rc + nc] = xb[iC * G::S2 + nc];
    }
  }
  static BTK_HD void step3_dft(cf* v) {
    BTK_UNROLL
    for (int rep = 0; rep < G::RepC; rep++) Dft<G::Rc, S>::run(v + rep * G::Rc);
  }

  // step 2 (only when Rb > 1): gather for pass B, radix-Rb, twiddle, scatter into exchange 2.
  // Reads complete before the caller's barrier; writes must come after it (same buffer is reused),
  // hence the split into step2_load / step2_store.
  static BTK_HD void step2_load(cf* v, int gl, const cf* xb, const cf* twb) {
    if (G::Rb > 1) {
      const float4 s01 = *reinterpret_cast<const float4*>(twb + gl * FT::TB);
      const cf q = twb[gl * FT::TB + 2];
      cf a = mk(s01.x, s01.y);                 // W^{nc ka} of the current rep
      const cf step = mk(s01.z, s01.w);
      BTK_UNROLL
      for (int rep = 0; rep < G::RepB; rep++) {
        const int iB = gl + G::L * rep;
        const int ka = iB / G::Rc, nc = iB % G::Rc;
        cf* p = v + rep * G::Rb;
        BTK_UNROLL
        for (int nb = 0; nb < G::Rb; nb++) p[nb] = xb[ka * G::S1 + nb * G::Rc + nc];
        Dft<G::Rb, S>::run(p);
        cf t = a;
        BTK_UNROLL
        for (int kb = 0; kb < G::Rb; kb++) {
          p[kb] = multw<S>(p[kb], t);
          if (kb + 1 < G::Rb) t = cmul(t, q);
        }
        if (rep + 1 < G::RepB) a = cmul(a, step);
      }
    }
  }
  static BTK_HD void step2_store(const cf* v, int gl, cf* xb) {
    if (G::Rb > 1) {
      BTK_UNROLL
      for (int rep = 0; rep < G::RepB; rep++) {
        const int iB = gl + G::L * rep;
        const int ka = iB / G::Rc, nc = iB % G::Rc;
        BTK_UNROLL
        for (int kb = 0; kb < G::Rb; kb++) xb[(ka + G::Ra * kb) * G::S2 + nc] = v[rep * G::Rb + kb];
      }
    }
  }

  // ---- transposed flow for ASYM plans (spectrum order in, time-domain order out):
  //   n = nc + Rc na, k = ka + Ra kc:  x[n] = sum_ka W_Ra^{S na ka} W_M^{S nc ka} sum_kc X[ka + Ra kc] W_Rc^{S nc kc}
  // inv_step1: radix-Rc over kc (lane holds ka = gl + L rep), twiddle W_M^{S nc ka}, scatter to xb[ka * S2 + nc];
  // inv_step3: lane nc = gl gathers all ka and runs radix-Ra: value na is sample gl + Rc na = index_of(gl, na).
  // The twiddle seed W_M^{ka} is twa[gl] (ka < L) times the constant W_M^{L rep}.
  template <int PP>
  static BTK_HD void inv_step1_multi(cf* v, int gl, cf* xb, const cf* twa) {
    static_assert(!G::ASYM || G::Rc <= 16, "product tree of the pass twiddles");
    static_assert(!G::ASYM || (G::RepC <= 2 && (32 * G::L == M_ || 16 * G::L == M_)), "W_M^L is hard-wired");
    BTK_UNROLL
    for (int rep = 0; rep < G::RepC; rep++) {
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) Dft<G::Rc, S>::run(v + pp * G::V + rep * G::Rc);
      const int ka = gl + G::L * rep;
      cf w[G::Rc + 1];
      w[1] = twa[gl * FT::TA];
      // W_M^{L} with 2 pi L / M = pi / 16 (M = 32 L) or pi / 8 (M = 16 L), see the static_assert below
      if (rep > 0) w[1] = cmul(w[1], M_ == 32 * G::L ? mk(BTK_COS_PI_16, BTK_SIN_PI_16) : mk(BTK_COS_PI_8, BTK_SIN_PI_8));
      BTK_UNROLL
      for (int nc = 2; nc < G::Rc; nc++) {
        const int hi = nc >= 8 ? 8 : (nc >= 4 ? 4 : 2);
        w[nc] = (nc == hi) ? cmul(w[nc / 2], w[nc / 2]) : cmul(w[hi], w[nc - hi]);
      }
      BTK_UNROLL
      for (int pp = 0; pp < PP; pp++) {
        cf* p = v + pp * G::V + rep * G::Rc;
        cf* x = xb + pp * G::XBUF;
        BTK_UNROLL
        for (int nc = 1; nc < G::Rc; nc++) p[nc] = multw<S>(p[nc], w[nc]);
        BTK_UNROLL
        for (int nc = 0; nc < G::Rc; nc++) x[ka * G::S2 + nc] = p[nc];
      }
    }
  }
  static BTK_HD void inv_step3(cf* v, int gl, const cf* xb) {
    static_assert(!G::ASYM || G::RepA == 1, "one radix-Ra transform per lane");
    BTK_UNROLL
    for (int ka = 0; ka < G::Ra; ka++) v[ka] = xb[ka * G::S2 + gl];
    Dft<G::Ra, S>::run(v);
  }

  // step 3: gather for pass C and radix-Rc; result in the canonical register layout.
  static BTK_HD void step3(cf* v, int gl, const cf* xb) {
    BTK_UNROLL
    for (int rep = 0; rep < G::RepC; rep++) {
      const int iC = gl + G::L * rep;
      cf* p = v + rep * G::Rc;
      BTK_UNROLL
      for (int nc = 0; nc < G::Rc; nc++) p[nc] = xb[iC * G::S2 + nc];
      Dft<G::Rc, S>::run(p);
    }
  }
};

}  // namespace btk
