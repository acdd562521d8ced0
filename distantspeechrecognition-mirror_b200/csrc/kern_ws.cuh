// kern_ws.cuh -- __global__ wrapper, device context and launch dispatch of the warp-specialised fused chain
// (chain_ws.cuh).  Included by kern_ws_m<M>.cu with one transform size per translation unit.
//
// Roles inside a CTA of NT + 256 threads (one CTA per SM, launched with 128 registers per thread, setmaxnreg on every warpgroup):
//   threads [0, NT)              transform warps       up to 192 / 216 registers: polyphase, transforms, weight accumulate
//   threads [NT, NT + 128)       overlap-add warpgroup 104 / 56 registers (tensor-copy mode): takes the v frames of every
//                                iteration out of tensor memory and runs the synthesis polyphase + overlap-add + stores
//   threads [NT + 128, NT + 256) producer warpgroup    tensor-copy mode: ONE thread issues cp.async.bulk.tensor boxes and the
//                                weight rows (24 registers); register-load mode: four warps of 16-byte loads
// Synchronisation: mbarriers in shared memory (full / empty per stage, one for the tables, full / empty per tensor-memory
// slot), named barriers for the transform warps (1) and the overlap-add warps (2), cluster-scope mbarriers for the
// channel-split reduction.  Bulk asynchronous copies (cp.async.bulk, SASS UBLKCP) bring the tap / twiddle tables once per
// CTA and the weight rows of every stage; tcgen05.alloc / st / ld / dealloc (UTCATOMSWS, STTM, LDTM) carry the v frames.
// The CTA is persistent: it walks its share of the launch as segments (chain_ws.cuh::WsSegs).
#pragma once

#include <cooperative_groups.h>

#include "chain_ws.cuh"
#include "launch.h"

namespace btk {

namespace cg = cooperative_groups;

#define BTK_MAX_SMEM_WS (227 * 1024)

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ws_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WS_DONE;\n"
      "bra WS_WAIT;\n"
      "WS_DONE:\n"
      "}\n" ::"r"(ws_smem_u32(b)), "r"(parity)
      : "memory");
}
// the producer's wait: it is a stage ahead most of the time, so it backs off between polls instead of taking issue slots
// from the transform warps
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WS_BWAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WS_BDONE;\n"
      "nanosleep.u32 200;\n"
      "bra WS_BWAIT;\n"
      "WS_BDONE:\n"
      "}\n" ::"r"(ws_smem_u32(b)), "r"(parity)
      : "memory");
}
// arrive on the same barrier of CTA `rank` of the cluster (default semantics: release at CTA scope, no GPU-wide fence)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, int rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(ws_smem_u32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// contiguous global -> shared copy by the bulk-copy engine; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ws_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(ws_smem_u32(bar))
               : "memory");
}

// One box of the staged window by the tensor copy engine (SASS UTMALDG): 4 channels x `rows` time steps of the recording's
// [T][C] tensor, starting at (channel c0, time step t0), land as [time step][4] at dst and complete their bytes on the
// mbarrier.  Time steps before 0 or past T are outside the tensor: the engine fills zeros, which is exactly the bank's
// zero history / zero tail (modulated.cc:461-516).
__device__ __forceinline__ void tma_box_g2s(void* dst, const void* tmap, int c0, int t0, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   ws_smem_u32(dst)),
               "l"(tmap), "r"(c0), "r"(t0), "r"(ws_smem_u32(bar))
               : "memory");
}

// tensor memory as a parking area (32x32b shape: lane t of the warp <-> lane 32 (warp mod 4) + t, N consecutive columns)
template <int N> struct TmemIO;
#define BTK_TM_R8(a, o) "%" #a ", %" #a "+1"
template <> struct TmemIO<8> {
  static __device__ __forceinline__ void st(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
  }
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
  }
};
template <> struct TmemIO<16> {
  static __device__ __forceinline__ void st(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
                 "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
  }
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                   "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr)
                 : "memory");
  }
};
template <> struct TmemIO<32> {
  static __device__ __forceinline__ void st(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]),
        "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]),
        "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
  }
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]),
          "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
          "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]),
          "=f"(v[31])
        : "r"(taddr)
        : "memory");
  }
};
template <> struct TmemIO<64> {
  static __device__ __forceinline__ void st(uint32_t taddr, const float* v) { TmemIO<32>::st(taddr, v); TmemIO<32>::st(taddr + 32, v + 32); }
  static __device__ __forceinline__ void ld(uint32_t taddr, float* v) { TmemIO<32>::ld(taddr, v); TmemIO<32>::ld(taddr + 32, v + 32); }
};
#undef BTK_TM_R8

template <int M, int PP, int NT> struct DevCtxWS {
  ChainThreadState<M, PP> ts;
  uint64_t* bars;
  int csz, crank;
  unsigned cl_phase;
  template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x, ts); }
  __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
  __device__ __forceinline__ void syncwarp() { __syncwarp(); }
  template <class F> __device__ __forceinline__ void acquire(int stage, int parity, F) { mbar_wait(bars + WS_BAR_FULL + stage, parity); }
  __device__ __forceinline__ void release(int stage) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + WS_BAR_EMPTY + stage);
  }
  __device__ __forceinline__ void wait_tables() { mbar_wait(bars + WS_BAR_TABLES, 0); }
  // ---- hand-over of the v frames to the overlap-add warps through tensor memory (chain_ws.cuh)
  uint32_t tmem;        // base address of the CTA's allocation
  int itc;              // iterations this CTA has published (runs on across segments)
  bool syn;
  static constexpr int NVAL = 2 * FFTGeom<M>::V;
  __device__ __forceinline__ bool syn_ok(const ChainParams&) const { return syn; }
  template <class F> __device__ __forceinline__ void syn_begin_segment(F) {}          // the overlap-add warps zero their ring themselves
  __device__ __forceinline__ int v_slot() const { return itc & 1; }
  __device__ __forceinline__ void v_acquire(int slot) { mbar_wait(bars + WS_BAR_VEMPTY + slot, ((itc >> 1) & 1) ^ 1); }
  __device__ __forceinline__ void tmem_store(int slot, int tid, const float* vals) {
    const int warp = tid >> 5;
    TmemIO<NVAL>::st(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((slot * 2 + (warp >> 2)) * NVAL), vals);
  }
  // second channel of the lane's pair, parked per residue step between the rounds of a stage (polyphase_pairs2): columns
  // zb0 + (warp / 4) * zbw + step * npk of the warp's lane quarter
  int zb0, zbw;
  template <int NPKV> __device__ __forceinline__ void zb_park_n(int tid, int step, const float* pk) {
    const int warp = tid >> 5;
    TmemIO<NPKV>::st(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(zb0 + (warp >> 2) * zbw + step * NPKV), pk);
  }
  __device__ __forceinline__ void zb_parked() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  template <int PC> __device__ __forceinline__ void zb_fetch_n(int tid, int col0, float* vals) {
    const int warp = tid >> 5;
    TmemIO<PC>::ld(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(zb0 + (warp >> 2) * zbw + col0), vals);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  template <class F> __device__ __forceinline__ void v_publish(int slot, F) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + WS_BAR_VFULL + slot);
    itc++;
  }
  __device__ __forceinline__ int cl_rank() const { return crank; }
  // Rendezvous of the compute warps of the whole cluster: every warp arrives on the READY barrier of every rank and waits
  // on its own.  Plain (CTA-scope release) remote arrives: the rendezvous only orders this CTA's earlier shared-memory
  // READS against the peers' later writes, the data itself is synchronised by the transaction counts below.
  __device__ __forceinline__ void cl_ready() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
      for (int r = 0; r < csz; r++) mbar_arrive_remote(bars + WS_BAR_READY, r);
    mbar_wait(bars + WS_BAR_READY, cl_phase & 1u);
    cl_phase ^= 1u;
  }
  __device__ __forceinline__ void cl_expect(int k, unsigned bytes) {
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bars + WS_BAR_RX0 + k, bytes);
  }
  __device__ __forceinline__ void cl_send16(void* local_dst, int rank, float4 v, int k) {
    uint32_t daddr, baddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(daddr) : "r"(ws_smem_u32(local_dst)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(baddr) : "r"(ws_smem_u32(bars + WS_BAR_RX0 + k)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(daddr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(baddr)
                 : "memory");
  }
  __device__ __forceinline__ void cl_wait(int k) {
    mbar_wait(bars + WS_BAR_RX0 + k, (cl_phase >> (1 + k)) & 1u);
    cl_phase ^= 2u << k;
  }
};

// context of the overlap-add warps: threads [NT, NT + NST) of the CTA, their own named barrier
template <int M, int PP, int NT, int NST> struct DevCtxSyn {
  ChainThreadState<M, PP> ts;      // never touched (the overlap-add programs keep no transform state)
  uint64_t* bars;
  uint32_t tmem;
  int itc;
  static constexpr int NVAL = 2 * FFTGeom<M>::V;
  template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x - NT, ts); }
  __device__ __forceinline__ void sync() { asm volatile("bar.sync 2, %0;" ::"n"(NST) : "memory"); }
  __device__ __forceinline__ void syncwarp() { __syncwarp(); }
  __device__ __forceinline__ void v_wait(int slot) {
    mbar_wait(bars + WS_BAR_VFULL + slot, (itc >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // the NVAL floats lane (tid mod 32) of transform warp w parked; this warp is warp (w mod 4) of its warpgroup
  template <int N> __device__ __forceinline__ void tmem_load(int slot, int w, int col0, int tid, float* vals) {
    TmemIO<N>::ld(tmem + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((slot * 2 + (w >> 2)) * NVAL + col0), vals);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  __device__ __forceinline__ void v_release(int slot) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bars + WS_BAR_VEMPTY + slot);
  }
};

// Registers per thread of the two roles.  The CTA is launched with LR registers per thread (what __launch_bounds__ of
// NT + 128 threads, one CTA per SM, allows); setmaxnreg.inc of the compute warpgroups can only take what setmaxnreg.dec of
// the producer warpgroup has returned to the CTA's pool, so 128 (LR - RP) >= NT (RC - LR) must hold or the compute warps
// spin in the allocation for ever.  RC covers the largest register program of each size (cuobjdump: 187 registers for
// M <= 256 with two frame pairs per warp, 209 for M = 512); the producers keep the rest for loads in flight.
template <int M, int NT> struct WsRegs {
  // CTA = NT transform threads + the overlap-add warpgroup + the producer warpgroup
  static constexpr int NTH = NT + 256;
  static constexpr bool split = true;
  static constexpr int LR = 65536 / NTH / 8 * 8;
  static constexpr int RC = M >= 512 ? 216 : 192;
  static constexpr int RMIN = 24;                                // a warpgroup that only polls a barrier or leaves at once
  // what the other working warpgroup gets: the pool minus the transform warps and the idle warpgroup, below the launch value
  static constexpr int RREST_RAW = (65536 - NT * RC - 128 * RMIN) / 128 / 8 * 8;
  static constexpr int RREST = RREST_RAW < LR ? RREST_RAW : LR - 8;
  static constexpr int RP = RREST;                               // producer warps (register-load mode)
  static constexpr int RS = RREST;                               // overlap-add warps (tensor-copy mode)
  static_assert(RC > LR && RREST >= 40 && NT * RC + 128 * RREST + 128 * RMIN <= 65536, "register pool balance");
};
// loads in flight per producer thread follow its register budget (a task is LV 16-byte loads)
template <int M, int NT, int LV> struct WsProd {
  static constexpr int RP = WsRegs<M, NT>::RP;
#ifndef BTK_WS_LOADS_HI
#define BTK_WS_LOADS_HI 8        // measured: 8 loads in flight beat 12 (cfg2 0.3479 -> 0.3419 ms) and 4 (0.3854 ms): a gentler producer disturbs the transform warps less
#endif
#ifndef BTK_WS_LOADS_MID
#define BTK_WS_LOADS_MID 8
#endif
  static constexpr int LOADS = RP >= 112 ? BTK_WS_LOADS_HI : (RP >= 64 ? BTK_WS_LOADS_MID : 4);
  static constexpr int TB = LOADS / LV;
};

template <int M, int R, int MT, int PP, bool SYNT>
__global__ void __launch_bounds__(WsCfg<M, R, MT, PP>::NT + 256, 1) btk_chain_ws_kernel(const ChainParams p) {
  typedef WsCfg<M, R, MT, PP> K;
  typedef WsRegs<M, K::NT> RG;
  extern __shared__ __align__(128) unsigned char smem[];
  const int m = MT > 0 ? MT : p.m;
  const WsSmem S = ws_smem_layout<M, R, PP>(m);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S.bars);
  // SYNT: overlap-add warps + tensor memory hand-over (chain_ws.cuh::ws_syn_mode, decided by the launcher): tensor-copy
  // producer, no cluster.  The other build carries the register-load producer, the cluster exchange and the synthesis
  // side on the transform warps.
  const int csz = SYNT ? 1 : (p.cluster > 1 ? p.cluster : 1);
  const int crank = csz > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int cta = (int)blockIdx.x / csz, ncta = (int)gridDim.x / csz;      // this CTA's (cluster's) share of the launch
  constexpr bool syn = SYNT;
  const bool tma = SYNT || p.tmaps != nullptr;
  __shared__ uint32_t s_tmem;

  if (threadIdx.x == 0) {
    for (int s = 0; s < K::NS; s++) {
      mbar_init(bars + WS_BAR_FULL + s, tma ? 1 : K::NPT);
      mbar_init(bars + WS_BAR_EMPTY + s, K::NW);
    }
    mbar_init(bars + WS_BAR_TABLES, 1);
    mbar_init(bars + WS_BAR_READY, csz * K::NW);
    mbar_init(bars + WS_BAR_RX0, 1);
    mbar_init(bars + WS_BAR_RX1, 1);
    for (int s = 0; s < 2; s++) {
      mbar_init(bars + WS_BAR_VFULL + s, K::NW);
      mbar_init(bars + WS_BAR_VEMPTY + s, K::NSY);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (syn && (int)threadIdx.x / 32 == K::NT / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ws_smem_u32(&s_tmem)), "n"(K::TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (csz > 1) cg::this_cluster().sync();      // nobody arrives on a remote barrier before it is initialised

  if (threadIdx.x >= K::NT && threadIdx.x < K::NT + K::NST) {
    // ------------------------------------------------------------------ overlap-add warpgroup
    if (!syn) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::RMIN));
    } else {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::RS));
      DevCtxSyn<M, PP, K::NT, K::NST> sctx;
      sctx.bars = bars; sctx.tmem = s_tmem; sctx.itc = 0;
      WsSegs segs(p, cta, ncta);
      WorkItem wk;
      WsSynState st;
      while (segs.next(wk)) {
        const RecDesc rec = p.recs[wk.rec];
        const int n_it = (wk.nj + S.L.H + K::W - 1) / K::W;
        chain_ws_synth_begin<M, R, MT, PP>(sctx, smem, S, st);
        for (int it = 0; it < n_it; it++, sctx.itc++) {
          const int slot = sctx.itc & 1;
          sctx.v_wait(slot);
          chain_ws_synth_iter<M, R, MT, PP>(sctx, p, smem, S, wk, rec, it, slot, st);
        }
      }
      // the allocation goes back when nobody touches tensor memory any more
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 3, %0;" ::"n"(K::NT + K::NST) : "memory");
      if ((int)threadIdx.x / 32 == K::NT / 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sctx.tmem), "n"(K::TM_COLS) : "memory");
    }
  } else if (threadIdx.x >= K::NT + K::NST) {
    // ------------------------------------------------------------------ producer warpgroup
    if (tma) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::RMIN));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::RP));
    const int ptid = (int)threadIdx.x - K::NT - K::NST;
    const ChainSmem& L = S.L;
    if (ptid == 0) {
      typedef FFTTables<M> FT;
      const uint32_t b_taps = K::D * L.TS * 4, b_twa = FT::TWA_WORDS * 8, b_twb = FT::TWB_WORDS * 8;
      mbar_arrive_expect_tx(bars + WS_BAR_TABLES, b_taps + b_twa + b_twb);
      bulk_g2s(smem + L.taps, p.taps_h, b_taps, bars + WS_BAR_TABLES);
      bulk_g2s(smem + L.twa, p.twa, b_twa, bars + WS_BAR_TABLES);
      if (b_twb) bulk_g2s(smem + L.twb, p.twb, b_twb, bars + WS_BAR_TABLES);
    }
    const int N = M * m;
    WsSegs segs(p, cta, ncta);
    WorkItem wk;
    int g = 0;                                   // stages filled so far by this CTA, across its segments
    if (tma) {
      // ---- tensor-copy producer: ONE thread.  Per stage it waits for the stage, announces the bytes and issues the
      // boxes of the window (NB D / rows of them) plus the weight rows; the copies run in the async proxy, no register
      // and no load/store-unit slot of this SM is involved, and the transform warps see the window in the layout of the
      // input ([time step][4 channels], WsCfg::RAW).
      if (ptid == 0) while (segs.next(wk)) {
        const WsWalk walk = ws_walk<K>(p, wk, L.H, csz, crank);
        const cf* wts = p.wts + (long long)wk.rec * p.wts_stride;
        const char* tmap = reinterpret_cast<const char*>(p.tmaps) + (size_t)wk.rec * 128;
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
        const int rows = p.tma_rows, nbox = (L.NB * K::D) / rows;
        const uint32_t stage_tx = (uint32_t)(L.NB * K::D * K::CG * 4 + K::CG * M * 8);
        for (int it = 0; it < walk.n_it; it++) {
          const int t_lo = (int)ws_window_start<K>(walk, it, p.laN, N);
          for (int cgi = 0; cgi < walk.ncg; cgi++, g++) {
            const int st = g % K::NS;
            unsigned char* stage = smem + S.stage0 + st * S.stage_bytes;
            const int cg0 = (walk.cg_base + cgi) * K::CG;
            mbar_wait_backoff(bars + WS_BAR_EMPTY + st, ((g / K::NS) & 1) ^ 1);
            mbar_arrive_expect_tx(bars + WS_BAR_FULL + st, stage_tx);
            bulk_g2s(stage + S.wts_off, wts + (long long)cg0 * M, K::CG * M * 8, bars + WS_BAR_FULL + st);
            for (int b = 0; b < nbox; b++)
              tma_box_g2s(stage + (size_t)b * rows * K::CG * 4, tmap, cg0, t_lo + b * rows, bars + WS_BAR_FULL + st);
          }
        }
      }
    } else if constexpr (!SYNT) while (segs.next(wk)) {
    constexpr int TB = WsProd<M, K::NT, K::LV>::TB;
    const RecDesc rec = p.recs[wk.rec];
    const WsWalk walk = ws_walk<K>(p, wk, L.H, csz, crank);
    const float* pcm = p.pcm + rec.pcm_off;
    const cf* wts = p.wts + (long long)wk.rec * p.wts_stride;
    const bool vec4 = (p.C % 4 == 0) && (rec.pcm_off % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.pcm) & 15) == 0);
    for (int it = 0; it < walk.n_it; it++) {
      const long long t_lo = ws_window_start<K>(walk, it, p.laN, N);
#ifdef BTK_WS_L2PF          // A/B, OFF by default: measured neutral (cfg2 0.3465 -> 0.3468 ms, cfg3 0.8292 -> 0.8308 ms)
      // The rows the NEXT iteration adds to the window (W D time steps, every channel: one contiguous run of the
      // interleaved recording) are asked into L2 now, by the bulk-prefetch engine: no registers, no scoreboard, any number
      // in flight.  The register loads of the next iteration's stages then run at L2 latency instead of DRAM latency
      // (a stage is 3-4 dependent batches of loads).  Skipped for many-channel inputs, where the windows in flight
      // already fill L2 (p.no_prefetch).
      if (!p.no_prefetch && it + 1 < walk.n_it && ptid < 32) {
        long long lo = t_lo + (long long)L.NB * K::D, hi = lo + (long long)K::W * K::D;
        if (lo < 0) lo = 0;
        if (hi > rec.T) hi = rec.T;
        const long long b0 = lo * p.C * 4, b1 = hi * p.C * 4;                 // byte range inside the recording
        constexpr long long CH = 4096;
        const char* base = reinterpret_cast<const char*>(pcm);
        for (long long o = (b0 & ~15ll) + (long long)ptid * CH; o < b1; o += 32 * CH) {
          long long n = b1 - o < CH ? b1 - o : CH;
          n = (n + 15) & ~15ll;
          if ((reinterpret_cast<uintptr_t>(base + o) & 15) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + o), "r"((uint32_t)n) : "memory");
        }
      }
#endif
      for (int cgi = 0; cgi < walk.ncg; cgi++, g++) {
        const int st = g % K::NS;
        unsigned char* stage = smem + S.stage0 + st * S.stage_bytes;
        const int cg0 = (walk.cg_base + cgi) * K::CG;
        const WsFill<K> fill(L, reinterpret_cast<float*>(stage), pcm, p.C, rec.T, t_lo, cg0, vec4, TB);
        float x[TB][K::LV][K::CG];
#ifdef BTK_WS_EARLY         // A/B, OFF by default: cfg2 0.3465 -> 0.3447 ms, but cfg3 0.8292 -> 0.8357 ms and cfg4 1.877 -> 1.942 ms
        fill.template load<TB>(ptid, 0, x);          // in flight while the stage is still with the transform warps
#endif
#ifdef BTK_WS_NO_BACKOFF
        mbar_wait(bars + WS_BAR_EMPTY + st, ((g / K::NS) & 1) ^ 1);
#else
        mbar_wait_backoff(bars + WS_BAR_EMPTY + st, ((g / K::NS) & 1) ^ 1);
#endif
        if (ptid == 0) {
          mbar_expect_tx(bars + WS_BAR_FULL + st, K::CG * M * 8);
          bulk_g2s(stage + S.wts_off, wts + (long long)cg0 * M, K::CG * M * 8, bars + WS_BAR_FULL + st);
        }
#ifndef BTK_EXP_NOFILL      // experiment: the compute side alone (stages handed over unfilled; results are garbage)
        for (int b = 0; b < fill.nbatch; b++) {
#ifdef BTK_WS_EARLY
          if (b > 0)
#endif
            fill.template load<TB>(ptid, b, x);
          fill.template store<TB>(ptid, b, x);
        }
#endif
        mbar_arrive(bars + WS_BAR_FULL + st);
      }
    }
    }
  } else {
    // ------------------------------------------------------------------ compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RG::RC));
    DevCtxWS<M, PP, K::NT> ctx;
    ctx.bars = bars; ctx.csz = csz; ctx.crank = crank; ctx.cl_phase = 0;
    ctx.tmem = syn ? s_tmem : 0u; ctx.itc = 0; ctx.syn = syn; ctx.zb0 = K::ZB0; ctx.zbw = K::ZBW;
    WsSegs segs(p, cta, ncta);
    WorkItem wk;
    int g = 0;                                   // stages consumed so far by this CTA, across its segments
    while (segs.next(wk)) chain_ws_compute<M, R, MT, PP, DevCtxWS<M, PP, K::NT>, SYNT ? 1 : 0>(ctx, p, smem, wk, p.recs[wk.rec], g);
    if (syn) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 3, %0;" ::"n"(K::NT + K::NST) : "memory");
    }
  }
  // no CTA of a cluster leaves while a peer may still write into its shared memory
  if (csz > 1) cg::this_cluster().sync();
}

// two frame pairs per warp for M <= 256 where the layout fits, else one; 0 = the shape does not fit at all
template <int M, int R> static int chain_ws_pp(int m) {
  if (M <= 256 && ws_smem_layout<M, R, 2>(m).total <= BTK_MAX_SMEM_WS) return 2;
  if (ws_smem_layout<M, R, 1>(m).total <= BTK_MAX_SMEM_WS) return 1;
  return 0;
}

template <int M, int R, int MT, int PP>
static cudaError_t launch_ws_one(const ChainParams& p, int n_work, cudaStream_t st) {
  typedef WsCfg<M, R, MT, PP> K;
  const WsSmem S = ws_smem_layout<M, R, PP>(p.m);
  const int smem = S.total;
  const int csz = p.cluster > 1 ? p.cluster : 1;
  // two builds of the kernel: with the overlap-add warpgroup (tensor-copy producer, no cluster) and without
  const bool syn = ws_syn_mode<K>(S, p.tmaps != nullptr && !(p.no_syn & 1), csz);
  auto kern = btk_chain_ws_kernel<M, R, MT, PP, false>;
  if constexpr (K::NG == 2 && MT > 0) {
    if (syn) kern = btk_chain_ws_kernel<M, R, MT, PP, true>;
  }
  if (csz > 1 && !ws_cluster_ok<K>(csz)) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_work * csz), 1, 1);
  cfg.blockDim = dim3(K::NT + 256, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.item_begin) {
    // persistent schedule: as many CTAs (clusters) as are resident at once -- one per SM; with a cluster size the GPC
    // shapes decide (not every SM pairs up), so ask the occupancy calculator rather than dividing the SM count
    static int slots[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (csz > 8) return cudaErrorInvalidValue;
    if (!slots[csz]) {
      int nc = 0;
      e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
      if (e != cudaSuccess) return e;
      slots[csz] = nc > 0 ? nc : 1;
    }
    int n_cta = n_work < slots[csz] ? n_work : slots[csz];
    if (p.cta_begin) {
      // host-balanced boundaries were made for p.cta_n CTAs, all resident at once; if this kernel cannot have that many,
      // split the items evenly instead
      if (p.cta_n > slots[csz]) {
        ChainParams q = p;
        q.cta_begin = nullptr; q.cta_n = 0;
        cfg.gridDim = dim3((unsigned)(n_cta * csz), 1, 1);
        return cudaLaunchKernelEx(&cfg, kern, q);
      }
      n_cta = p.cta_n;
    }
    cfg.gridDim = dim3((unsigned)(n_cta * csz), 1, 1);
  }
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int M, int R, int MT> struct WsFastOk { static constexpr bool value = MT * R <= 16; };
#define BTK_WS_MT(MTV) (WsFastOk<M, R, MTV>::value ? MTV : 0)

template <int M, int R>
static cudaError_t launch_chain_ws_r(const ChainParams& p, int n_work, cudaStream_t st) {
  const int pp = chain_ws_pp<M, R>(p.m);
  if (pp == 0) return cudaErrorInvalidValue;
  if (M <= 256 && pp == 2) {
    constexpr int PP = M <= 256 ? 2 : 1;
    if (p.m == 2 && WsFastOk<M, R, 2>::value) return launch_ws_one<M, R, BTK_WS_MT(2), PP>(p, n_work, st);
    if (p.m == 4 && WsFastOk<M, R, 4>::value) return launch_ws_one<M, R, BTK_WS_MT(4), PP>(p, n_work, st);
    return launch_ws_one<M, R, 0, PP>(p, n_work, st);
  }
  if (p.m == 2 && WsFastOk<M, R, 2>::value) return launch_ws_one<M, R, BTK_WS_MT(2), 1>(p, n_work, st);
  if (p.m == 4 && WsFastOk<M, R, 4>::value) return launch_ws_one<M, R, BTK_WS_MT(4), 1>(p, n_work, st);
  return launch_ws_one<M, R, 0, 1>(p, n_work, st);
}

template <int M, int R> static int chain_ws_w(int m) {
  const int pp = chain_ws_pp<M, R>(m);
  if (pp == 0) return -1;
  return pp == 2 ? WsCfg<M, R, 0, 2>::W : WsCfg<M, R, 0, 1>::W;
}
template <int M, int R> static bool chain_ws_cluster_ok(int m, int S) {
  const int pp = chain_ws_pp<M, R>(m);
  if (pp == 0) return false;
  return pp == 2 ? ws_cluster_ok<WsCfg<M, R, 0, 2> >(S) : ws_cluster_ok<WsCfg<M, R, 0, 1> >(S);
}

}  // namespace btk

#define BTK_DEFINE_WS_LAUNCHERS(MM)                                                                                \
  namespace btk {                                                                                                  \
  cudaError_t launch_chain_ws_m##MM(int R, const ChainParams& p, int n_work, cudaStream_t st) {                    \
    switch (R) {                                                                                                   \
      case 1: return launch_chain_ws_r<MM, 1>(p, n_work, st);                                                      \
      case 2: return launch_chain_ws_r<MM, 2>(p, n_work, st);                                                      \
      case 4: return launch_chain_ws_r<MM, 4>(p, n_work, st);                                                      \
      case 8: return launch_chain_ws_r<MM, 8>(p, n_work, st);                                                      \
    }                                                                                                              \
    return cudaErrorInvalidValue;                                                                                  \
  }                                                                                                                \
  int chain_ws_frames_per_iter_m##MM(int R, int m) {                                                               \
    switch (R) {                                                                                                   \
      case 1: return chain_ws_w<MM, 1>(m);                                                                         \
      case 2: return chain_ws_w<MM, 2>(m);                                                                         \
      case 4: return chain_ws_w<MM, 4>(m);                                                                         \
      case 8: return chain_ws_w<MM, 8>(m);                                                                         \
    }                                                                                                              \
    return -1;                                                                                                     \
  }                                                                                                                \
  bool chain_ws_cluster_ok_m##MM(int R, int m, int S) {                                                            \
    switch (R) {                                                                                                   \
      case 1: return chain_ws_cluster_ok<MM, 1>(m, S);                                                             \
      case 2: return chain_ws_cluster_ok<MM, 2>(m, S);                                                             \
      case 4: return chain_ws_cluster_ok<MM, 4>(m, S);                                                             \
      case 8: return chain_ws_cluster_ok<MM, 8>(m, S);                                                             \
    }                                                                                                              \
    return false;                                                                                                  \
  }                                                                                                                \
  }
