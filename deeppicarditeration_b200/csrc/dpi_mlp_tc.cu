// 16-bit (fp16 operands, fp16 accumulators in TMEM; fp32 accumulators for the chunked first layer of d > 128) tcgen05
// integral estimator for the Burgers-type equation `Cha` (u_k = MLP [1+d,128,128,128,128,1], d <= 1024).
//
// Same mathematics as dpi_mlp_simt.cu (picard/data.py:471-527,350-366,1226-1325; equations.py:297-302;
// solution.py:123-135) with the first-layer GEMM and the three hidden 128x128 GEMMs (value and tangent) on the
// 5th-generation tensor cores:
//
//   * one persistent CTA per SM.  A work unit is (collocation point, group of 16 tiles); a tile is 128 MC samples =
//     128 TMEM lanes.  TWO tiles are in flight per CTA, one per "slot": while the epilogue warps of one slot turn a
//     layer's accumulators into the next layer's operands, the tensor pipe runs the other slot's layer;
//   * shared memory: the hidden weights W2..W4 stay resident as fp16 K-major SWIZZLE_128B operand images (3 x 32 KB);
//     each slot owns an hV and an hT buffer (value rows / tangent rows of the next A operand, 2 x 32 KB).  For layer 1
//     the slot's hV holds the increments Z (A operand) and its hT receives the W1 image (B operand) by TMA bulk copy
//     from L2 for every tile -- that is what makes room for the second tile in 227 KB;
//   * TMEM: 512 columns = 2 slots x (value 128 + tangent 128) accumulator columns (fp16 accumulators keep one half per
//     32-bit cell and are read back two columns per register, tcgen05.ld ... .pack::16b: no conversions in the epilogue);
//   * 8 producer warps draw the Brownian increments with Philox4x32-10 + Box-Muller in registers, round them to fp16,
//     store the [128 x d] A-operand tile when the slot frees up, and KEEP them in registers: once the tile's Bismut
//     coefficients are known they accumulate G_j += coef_m Z_mj there, so increments never leave the SM;
//   * one thread issues tcgen05.mma for both slots in a fixed round-robin order (one issuing warp per slot was
//     measured 5 % slower), one thread issues the TMA copies and pulls the CTA's next work unit from a global counter;
//   * 8 epilogue warps (thread = sample row x column half) serve both slots in turn: they read the accumulators with
//     tcgen05.ld, apply bias/ELU/ELU' in packed fp16x2 arithmetic (two columns per instruction; its absolute error stays below the rounding of the
//     16-bit operands it produces) and write the next layer's A operands straight into the swizzled shared-memory
//     layout; the last layer is a dot product in the epilogue, followed by f, the Bismut coefficient and the
//     value-label partial sum in fp32.
//
// Partial sums go to a small scratch ([unit][row class][128], [unit][slot][lane quarter]) and a finish kernel adds
// them in a fixed order; the slot of a tile is its index inside the unit, so labels are bitwise independent of the
// launch geometry and of how points are sharded.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "dpi_common.cuh"
#include "dpi_kernels.h"
#include "dpi_tc_common.cuh"

namespace dpi {
namespace {
using namespace tc;

// ---- roles ---------------------------------------------------------------------------------------------------
// warps 0-7: epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = column half; both tile slots alternately),
// warps 8-15: increment producers, warp 16: MMA issuer (one thread), warp 17: TMA issuer (one thread) and unit
// hand-out, warps 18-19: idle.
// The register file is per SM sub-partition (512 registers per lane, warp w lives on sub-partition w % 4), so each
// sub-partition holds one warp of every warpgroup.  setmaxnreg only moves registers inside the CTA's own pool (what
// its warps have released), so the re-split must balance against the launch allocation of 96 per thread:
// MMA/TMA warpgroup 96 -> 32, epilogue stays at 96, producers 96 -> 128 (2 x 96 + 2 x 128 + 32 = 480 = 5 x 96;
// measured best of 80/144, 88/136, 96/128).
constexpr int kTcThreads = 640;
// Role -> warp ids.  The warp scheduler of a sub-partition prefers the HIGHER warp id among its eligible warps
// (B300_MICROARCH.md, multi-warp arbiter), and the epilogue is the critical path of a tile while a producer that polls
// a barrier is not: with DPI_TC_ROLE_SWAP the epilogue takes warps 8-15 and the producers warps 0-7.
#ifndef DPI_TC_ROLE_SWAP
#define DPI_TC_ROLE_SWAP 0
#endif
constexpr int kProdWarp0 = DPI_TC_ROLE_SWAP ? 0 : 8, kEpiWarp0 = DPI_TC_ROLE_SWAP ? 8 : 0, kMmaWarp = 16, kTmaWarp = 17;
#ifndef DPI_EPI_REGS
#define DPI_EPI_REGS 96
#endif
constexpr int kEpiRegs = DPI_EPI_REGS, kProdRegs = 224 - DPI_EPI_REGS, kMmaRegs = 32;
constexpr float kLog2e = 1.4426950408889634f;

// shared memory map (dynamic, 1024-byte aligned)
constexpr uint32_t kOffW = 0;                              // W2, W3, W4 operand images (3 x 32 KB), resident
constexpr uint32_t kOffSlot = 3 * kOperandBytes;           // tile slot s at + s * 64 KB: hV (32 KB) | hT (32 KB)
constexpr uint32_t kSlotBytes = 2 * kOperandBytes;
constexpr uint32_t kOffExch = kOffSlot + 2 * kSlotBytes;   // [128] float2: output-layer partials handed between the halves
constexpr uint32_t kOffCoef = kOffExch + 1024;              // [3][128] Bismut coefficients (ring over tiles)
constexpr uint32_t kOffBars = kOffCoef + 1536;             // 20 mbarriers
constexpr uint32_t kOffTmemPtr = kOffBars + 224;           // 26 mbarriers (8 B each) + padding
constexpr uint32_t kOffUnits = kOffTmemPtr + 16;           // [4] ring of unit ids handed out by the global counter
constexpr uint32_t kTcSmemBytes = kOffUnits + 16;          // 232,192 B <= 232,448 B

enum Bar : int {
  kBarW = 0,         // W2..W4 landed (TMA)
  kBarW1Full = 1,    // [2] W1 image landed in the slot's hT buffer (TMA, once per tile)
  kBarZFull = 3,     // [2] increments stored in the slot's hV buffer (8 producer warps)
  kBarSlotFree = 5,  // [2] layer-4 MMAs of the slot's tile have completed: hV / hT may be overwritten (tcgen05.commit)
  kBarAccFull = 7,   // [2] a layer's accumulators are complete (tcgen05.commit, 4 per tile)
  kBarAReady = 9,    // [2] an epilogue pass is done: next A operand written, accumulators read (8 epilogue warps)
  kBarCoef = 11,     // [3] Bismut coefficients of a tile written (4 epilogue warps)
  kBarUnit = 14,     // [4] the CTA's k-th unit id is in ring entry k % 4 (TMA warp)
  kBarRingFull = 18, // [4] MULTI: every chunk of tile j is in ring slot j % 4 of the global increment ring (8 producer warps)
  kBarRingFree = 22  // [4] MULTI: the Bismut sums of the tile in ring slot j % 4 are done (8 epilogue warps)
};

// d > 128: the Bismut sums G_j += coef_m Z_mj of a tile are taken by the EPILOGUE warps right after the tile's output
// pass (warp e owns first-layer chunk e), reading the tile's increments back from the global ring.  The producers used
// to do it two tiles later (DPI_MULTI_PROD_BISMUT restores that): ring write + read-back + sums sat on the producers'
// critical path, which is what bounds this kernel, while the epilogue warps idle for most of a tile.
#ifdef DPI_MULTI_PROD_BISMUT
constexpr bool kMultiEpiBismut = false;
#else
constexpr bool kMultiEpiBismut = true;
#endif

struct TcArgs {
  const void* W;  // 4 consecutive operand images: W1 (x columns), W2, W3, W4
  const float* tx;
  int64_t n_units;
  int64_t point_base;
  uint64_t seed;
  PhiloxKeys keys;  // round keys of `seed`
  const float* u_s;
  const float* zS;
  const float* fb;  // [n] control variate f(t,x)
  const float* c0;  // [n][128]
  float* partG;     // [units][16][128]
  float* partF;     // [units][8]
  uint4* zring;     // MULTI only: [grid][4 tiles][chunks][128 rows][16 octets] fp16x8 increments awaiting their coefficients
  int M_I;
  int units_per_point;
  int tiles_per_point;
  int ksteps1;  // ceil(d/16)
  int no;       // ceil(d/8) octets of dims (one Philox block, one 16-byte operand chunk each)
  int cls;      // row classes of the producer mapping: cls * no <= 256, cls <= kTcMaxClasses
  int nchunks;  // ceil(d / 128) first-layer K chunks (1 unless MULTI)
  int* unit_counter;  // next unit to hand out (zeroed before the launch)
  long long* trace;  // developer timeline (DPI_TC_TRACE builds only), else null
};

#ifdef DPI_TC_TRACE
// Low-overhead timeline of CTA 0: every traced thread owns a region of the buffer and a register counter, an event
// is one fire-and-forget 16-byte store {code = slot*100000 + (tile j % 1000)*100 + tag, clock64}.
constexpr int kTraceCap = 8192;
#define TRACE_DECL(region) long long* trace_p = (A.trace && blockIdx.x == 0) ? A.trace + (region) * 2 * kTraceCap : nullptr; int trace_n = 0
#define TRACE(role, slot, j, tag)                                                                         \
  do {                                                                                                    \
    if (trace_p && trace_n < kTraceCap) {                                                                 \
      *reinterpret_cast<longlong2*>(trace_p + 2 * trace_n) =                                              \
          make_longlong2((long long)(slot) * 100000 + (long long)((j) % 1000u) * 100 + (tag), clock64()); \
      ++trace_n;                                                                                          \
    }                                                                                                     \
  } while (0)
#else
#define TRACE_DECL(region)
#define TRACE(role, slot, j, tag)
#endif

// All activations are carried scaled by log2(e): with z' = log2e z the ELU pair needs one op less per element,
//   e = 2^min(z',0) = ELU'(z),   h' = log2e ELU(z) = max(z', e log2e - log2e),   tangent' = e * (log2e dz),
// and because every layer is linear in its input only the biases (x log2e), the first-layer constants and the
// output weights (x ln2) change -- the hidden weight matrices are used as they are (TcConsts holds the scaled
// vectors; see tc_pack_network).
//
// One epilogue pass over the 128 accumulator columns of this thread's row (TMEM lane), 8 columns at a time with the
// next chunk's tcgen05.ld in flight under the arithmetic of the current one.  Column indices are compile-time, so
// the per-column constants are read straight from the constant bank as instruction operands.
//   kPassFirst : z' = c0' + s' w_t' + a' acc, tangent' = e wsum' (constant first-layer tangent)  -> bf16 A operands
//   kPassHidden: z' = acc + b'[layer], tangent' = e acc_t                                          -> bf16 A operands
//   kPassOut   : same ELU, then the output layer as fp32 dot products with w_out: pu = w.h, pD = w.(e acc_t)
enum : int { kPassFirst = 0, kPassHidden = 1, kPassOut = 2 };
#ifndef DPI_EPI_UNROLL
#define DPI_EPI_UNROLL 4
#endif
constexpr int kEpiUnroll = DPI_EPI_UNROLL;  // chunks per loop iteration of an epilogue pass (even)

#ifndef DPI_EPI_CW
#define DPI_EPI_CW 16
#endif
#ifndef DPI_EPI_DB
#define DPI_EPI_DB 0
#endif
#ifndef DPI_EPI_POLY
#define DPI_EPI_POLY 0x1
#endif
// which of the four column pairs of every 8-column group take the FMA-pipe exponential (h2_ex2_poly) instead of MUFU
constexpr int kEpiPolyMask = DPI_EPI_POLY;
constexpr int kEpiCW = DPI_EPI_CW;       // accumulator columns per tcgen05.ld (8 or 16)
constexpr bool kEpiDB = DPI_EPI_DB != 0; // prefetch the next chunk under the arithmetic of the current one
// fp16 accumulators in TMEM (see make_idesc_f16_acc16): the epilogue reads packed pairs and skips the fp32 -> fp16
// conversions (2 of its ~8.5 instructions per column pair); the first layer of the chunked d > 128 path keeps fp32
// accumulators (its sums run over up to 1024 inputs)
#ifndef DPI_TC_ACC16
#define DPI_TC_ACC16 1
#endif
constexpr bool kAcc16 = DPI_TC_ACC16 != 0;
#ifndef DPI_TC_ACC16_FIRST
#define DPI_TC_ACC16_FIRST DPI_TC_ACC16
#endif
constexpr bool kAcc16First = DPI_TC_ACC16_FIRST != 0;  // the first layer too (its epilogue then runs in packed fp16)
static_assert(kEpiCW == 8 || kEpiCW == 16 || kEpiCW == 32, "chunk width");

template <int N>
__device__ __forceinline__ void tmem_ldN(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 8) tmem_ld8(taddr, r);
  else if constexpr (N == 16) tmem_ld16(taddr, r);
  else tmem_ld32(taddr, r);
}
template <int N>  // N registers = 2 N columns of fp16 accumulators
__device__ __forceinline__ void tmem_ldN_pack(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 4) tmem_ld4_pack(taddr, r);
  else if constexpr (N == 8) tmem_ld8_pack(taddr, r);
  else tmem_ld16_pack(taddr, r);
}

template <int MODE, bool ACC16 = false, int POLY = kEpiPolyMask>
__device__ __forceinline__ void epi_pass(const TcConsts& C, uint32_t tacc, uint8_t* hv_row, uint8_t* ht_row,
                                         uint32_t rx, int half, const float* __restrict__ c0_s, int layer, float s2,
                                         float a2, float& pu, float& pD) {
  // this warp's half of the row: columns [64 half, 64 half + 64); c0_s = the point's (log2e-scaled) first-layer
  // constants in global memory (512 B per point, read by every tile of the point: L1 hits)
  constexpr int NB = kEpiDB ? 2 : 1, NCH = 64 / kEpiCW, G = kEpiCW / 8;
  const int ch0 = half * NCH;
  constexpr int NR = ACC16 ? kEpiCW / 2 : kEpiCW;  // registers per chunk and accumulator row
  uint32_t v[NB][NR], tv[NB][NR];
  float4 cc[NB][2 * G];  // first pass: c0 of the current / next chunk (prefetched like the accumulators)
  const uint32_t a2h = pack_f16x2(a2, a2), s2h = pack_f16x2(s2, s2);
  auto fetch = [&](int ch, int b) {
    if constexpr (ACC16) {
      tmem_ldN_pack<NR>(tacc + (uint32_t)(ch * kEpiCW), v[b]);
      if (MODE != kPassFirst) tmem_ldN_pack<NR>(tacc + 128u + (uint32_t)(ch * kEpiCW), tv[b]);
    } else {
      tmem_ldN<kEpiCW>(tacc + (uint32_t)(ch * kEpiCW), v[b]);
      if (MODE != kPassFirst) tmem_ldN<kEpiCW>(tacc + 128u + (uint32_t)(ch * kEpiCW), tv[b]);
    }
    if (MODE == kPassFirst) {
#pragma unroll
      for (int i = 0; i < 2 * G; ++i) cc[b][i] = __ldg(reinterpret_cast<const float4*>(&c0_s[ch * kEpiCW + 4 * i]));
    }
  };
  if (kEpiDB) fetch(ch0, 0);
  const uint32_t* bias = C.b16[layer];
  const uint32_t kL = h2_const(kLog2e, kLog2e), kNegL = h2_const(-kLog2e, -kLog2e), kZero = 0u;
  uint32_t accU = 0u, accD = 0u;  // output pass: fp16x2 partial dot products, flushed to fp32 every 16 columns
  // rolled a few chunks at a time: the pass is a short loop that stays in the instruction cache; per-column
  // constants are then fetched with indexed constant loads
#pragma unroll kEpiUnroll
  for (int ci = 0; ci < NCH; ++ci) {
    const int ch = ch0 + ci;
    const int b = kEpiDB ? (ci & 1) : 0;
    if (!kEpiDB) fetch(ch, 0);
    tmem_ld_wait();
    if (kEpiDB && ci + 1 < NCH) fetch(ch + 1, (ci + 1) & 1);
#pragma unroll
    for (int g = 0; g < G; ++g) {  // groups of 8 columns = one 16-byte operand chunk of the value and tangent rows
      const uint32_t* vv = &v[b][(ACC16 ? 4 : 8) * g];
      const uint32_t* tt = &tv[b][(ACC16 ? 4 : 8) * g];
      const int c8 = ch * G + g;   // 16-byte chunk index 0..15
      uint32_t hv[4], ht[4];
#pragma unroll
      for (int k2 = 0; k2 < 8; k2 += 2) {  // two columns per packed instruction
        const int col = c8 * 8 + k2;
        uint32_t z, tg;
        if (MODE == kPassFirst) {
          const float4 c0 = cc[b][2 * g + k2 / 4];
          const float c0a = (k2 & 2) ? c0.z : c0.x, c0b = (k2 & 2) ? c0.w : c0.y;
          if constexpr (ACC16)  // packed: z' = a' acc + (s' w_t' + c0')
            z = h2_fma(vv[k2 / 2], a2h, h2_fma(s2h, C.wt16[col / 2], pack_f16x2(c0a, c0b)));
          else
            z = pack_f16x2(fmaf(a2, __uint_as_float(vv[k2]), fmaf(s2, C.wt[col], c0a)),
                           fmaf(a2, __uint_as_float(vv[k2 + 1]), fmaf(s2, C.wt[col + 1], c0b)));
          tg = C.wsum16[col / 2];
        } else if constexpr (ACC16) {
          z = h2_add(vv[k2 / 2], MODE == kPassHidden ? bias[col / 2] : C.b16[2][col / 2]);
          tg = tt[k2 / 2];
        } else {
          z = h2_add(pack_f16x2(__uint_as_float(vv[k2]), __uint_as_float(vv[k2 + 1])),
                     MODE == kPassHidden ? bias[col / 2] : C.b16[2][col / 2]);
          tg = pack_f16x2(__uint_as_float(tt[k2]), __uint_as_float(tt[k2 + 1]));
        }
        const uint32_t zneg = h2_min(z, kZero);
        const uint32_t e = ((POLY >> (k2 / 2)) & 1) ? h2_ex2_poly(zneg) : h2_ex2(zneg);  // ELU'(z) = 2^min(z',0)
        const uint32_t h = h2_max(z, h2_fma(e, kL, kNegL));    // log2e ELU(z)
        const uint32_t dh = h2_mul(e, tg);
        if (MODE == kPassOut) {
          accU = h2_fma(C.wout16[col / 2], h, accU);
          accD = h2_fma(C.wout16[col / 2], dh, accD);
        } else {
          hv[k2 / 2] = h;
          ht[k2 / 2] = dh;
        }
      }
      if (MODE == kPassOut) {
        if ((G == 1 ? ci : g) & 1) {
          const float2 fu = h2_unpack(accU), fd = h2_unpack(accD);
          pu += fu.x + fu.y;
          pD += fd.x + fd.y;
          accU = accD = 0u;
        }
      } else {
        // 16-byte chunk c8 of this row: K panel (c8 >> 3), chunk index XOR-swizzled with (row & 7)
        const uint32_t off = (uint32_t)(c8 >> 3) * kPanelBytes + (((uint32_t)(c8 & 7) << 4) ^ rx);
        *reinterpret_cast<uint4*>(hv_row + off) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
        *reinterpret_cast<uint4*>(ht_row + off) = make_uint4(ht[0], ht[1], ht[2], ht[3]);
      }
    }
  }
}

// Units are handed out dynamically: SMs do not run at exactly the same speed (3-8 % spread between the fastest and
// the slowest CTA of a statically partitioned launch), so each CTA pulls its next unit from a global counter.  The
// TMA warp -- the role that runs ahead of all others -- fetches one unit ahead and publishes the CTA's k-th unit id
// in a 4-entry shared-memory ring; every role walks the same sequence k = 0, 1, ... and stops at the first id
// >= n_units.  All roles are within one unit of each other, so an entry is never overwritten while it is still
// read, and a waiter can never be two phases behind.  Partial sums are indexed by the unit id: which CTA computes
// a unit changes nothing in the result.
struct UnitFeed {
  const volatile int* ring;
  uint32_t bar0;
  __device__ __forceinline__ int64_t get(uint32_t k) const {
    mbar_wait(bar0 + 8u * (k & 3u), (k >> 2) & 1u, 50);
    return (int64_t)ring[k & 3u];
  }
};

// flat cursor over this CTA's tiles: unit u = the CTA's k-th unit, tile tl inside the unit; slot = tl & 1, so the
// slot of a tile (and with it the order of every floating-point sum) does not depend on the launch geometry
struct TileCursor {
  int64_t u;
  int tl, tiles;
  uint32_t k;
  __device__ __forceinline__ void init(const TcArgs& A, const UnitFeed& F) {
    k = 0;
    u = F.get(0);
    tl = 0;
    tiles = (u < A.n_units) ? tiles_of(A, u) : 0;
  }
  __device__ __forceinline__ static int tiles_of(const TcArgs& A, int64_t uu) {
    const int g = (int)(uu % A.units_per_point);
    return min(kTcGroupTiles, A.tiles_per_point - g * kTcGroupTiles);
  }
  __device__ __forceinline__ bool valid(const TcArgs& A) const { return u < A.n_units; }
  __device__ __forceinline__ void next(const TcArgs& A, const UnitFeed& F) {
    if (++tl == tiles) {
      u = F.get(++k);
      tl = 0;
      tiles = (u < A.n_units) ? tiles_of(A, u) : 0;
    }
  }
};

// MULTI: 128 < d <= 1024.  The first layer is accumulated over ceil(d/128) chunks of 128 input dims: per chunk the
// producers store the increments of those dims into the slot's hV, the TMA warp streams the matching W1 image into
// hT, the MMA issuer adds the chunk's eight k-steps to the value accumulators and commits kBarSlotFree so that the
// next chunk may overwrite the buffers.  The increments of a tile no longer fit in registers (8 x as many), so
// they also go (fp16) to a per-CTA ring in global memory and are read back for the Bismut sums once the coefficients
// are known, warp w taking chunk w.
template <int ITEMS, bool DET, bool MULTI>
__global__ void __launch_bounds__(kTcThreads, 1)
cha_tc_kernel(const __grid_constant__ ProblemDev P, const __grid_constant__ TcConsts C,
              const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.nx;
  const int UP = A.units_per_point, M_I = A.M_I;
  auto bar = [&](int i) -> uint32_t { return sbase + kOffBars + 8u * (uint32_t)i; };
  const UnitFeed feed{reinterpret_cast<const volatile int*>(smem + kOffUnits), sbase + kOffBars + 8u * (uint32_t)kBarUnit};

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {
      printf("dpi tc: dynamic shared memory base 0x%x is not 1024-byte aligned\n", sbase);
      __trap();
    }
    mbar_init(bar(kBarW), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarW1Full + s), 1);
      mbar_init(bar(kBarZFull + s), 8);
      mbar_init(bar(kBarSlotFree + s), 1);
      mbar_init(bar(kBarAccFull + s), 1);
      mbar_init(bar(kBarAReady + s), 8);
    }
    for (int b = 0; b < 3; ++b) mbar_init(bar(kBarCoef + b), 4);
    for (int b = 0; b < 4; ++b) mbar_init(bar(kBarUnit + b), 1);
    for (int b = 0; b < 4; ++b) {
      mbar_init(bar(kBarRingFull + b), 8);
      mbar_init(bar(kBarRingFree + b), 8);
    }
    fence_barrier_init();
  }
  // The layer-1 MMA reads K = 16 ceil(d/16) columns of the slot's hV buffer; the producers write the first d of them
  // and the padded weight columns are zero, so whatever the buffer holds there only has to be finite: zero it once.
  for (int i = threadIdx.x; i < (int)(2 * kSlotBytes / 16); i += kTcThreads)
    reinterpret_cast<uint4*>(smem + kOffSlot)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == kMmaWarp) tmem_alloc(sbase + kOffTmemPtr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);
#ifdef DPI_TC_TRACE
  unsigned long long cta_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_t0));
#endif
  if (tmem != 0u) {  // all 512 columns of the SM belong to this CTA: the MMA issuer relies on base column 0
    if (threadIdx.x == 0) printf("dpi tc: unexpected TMEM base 0x%x\n", tmem);
    __trap();
  }

  if (warp >= kMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMmaRegs));
    if (warp == kMmaWarp) {
      // =================================== MMA issuer ===================================
      // The whole warp runs this (uniform) control flow; the elected lane issues tcgen05.mma / commit.
      const uint32_t leader = elect_one();
      TRACE_DECL(0);
#ifdef DPI_TC_TRACE
      if (lane != 0) trace_p = nullptr;
#endif
      mbar_wait(bar(kBarW), 0, 0);
      // accumulator type per layer: fp16 (kAcc16) except the chunked first layer of the d > 128 path
      constexpr uint32_t idesc = kAcc16 ? make_idesc_f16_acc16(128, 128) : make_idesc_f16(128, 128);
      constexpr uint32_t idesc1 = (kAcc16First && !MULTI) ? make_idesc_f16_acc16(128, 128) : make_idesc_f16(128, 128);
      TileCursor cur;
      cur.init(A, feed);
      int lay0 = 0, lay1 = 0, slt0 = 0, slt1 = 0;  // two in-flight tiles: next layer to issue (0 = none) and slot
      uint32_t ph_z = 0, ph_w1 = 0, ph_ar = 0, used = 0;  // bit s = slot s
      auto issue = [&](int s, int l) {
        TRACE(1, s, 0, l * 10 + 0);
        // lane-0 broadcasts mark these warp-uniform values as uniform for the compiler (uniform registers / datapath)
        s = __shfl_sync(0xffffffffu, s, 0);
        l = __shfl_sync(0xffffffffu, l, 0);
        const uint32_t acc = (uint32_t)s * 256u;  // TMEM base is column 0 (checked at allocation)
        const uint32_t slot = sbase + kOffSlot + (uint32_t)s * kSlotBytes;
        const uint64_t dV = make_desc_sw128(slot), dT = make_desc_sw128(slot + kOperandBytes);
        const uint32_t bit = 1u << s;
        if (l == 1) {
          // accumulators free (output pass of the slot's previous tile done), increments stored, W1 landed in hT
          if (used & bit) {
            mbar_wait(bar(kBarAReady + s), (ph_ar >> s) & 1u, 10);
            ph_ar ^= bit;
          }
          used |= bit;
          if constexpr (MULTI) {
            const int nch = A.nchunks;
            for (int cch = 0; cch < nch; ++cch) {
              mbar_wait(bar(kBarZFull + s), (ph_z >> s) & 1u, 11);
              ph_z ^= bit;
              mbar_wait(bar(kBarW1Full + s), (ph_w1 >> s) & 1u, 12);
              ph_w1 ^= bit;
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t o = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
                umma_bf16_p(acc, dV + o, dT + o, idesc1, (uint32_t)((cch | ks) != 0), leader);
              }
              if (cch + 1 < nch) tc_commit_p(bar(kBarSlotFree + s), leader);  // hV / hT may take the next chunk
            }
          } else {
            mbar_wait(bar(kBarZFull + s), (ph_z >> s) & 1u, 11);
            ph_z ^= bit;
            mbar_wait(bar(kBarW1Full + s), (ph_w1 >> s) & 1u, 12);
            ph_w1 ^= bit;
            tc_fence_after();
            TRACE(1, s, 0, l * 10 + 1);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ks < A.ksteps1) {
                const uint64_t o = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
                umma_bf16_p(acc, dV + o, dT + o, idesc1, ks > 0, leader);
              }
            }
          }
          tc_commit_p(bar(kBarAccFull + s), leader);
          TRACE(1, s, 0, l * 10 + 2);
        } else {
          mbar_wait(bar(kBarAReady + s), (ph_ar >> s) & 1u, 13);
          ph_ar ^= bit;
          tc_fence_after();
          TRACE(1, s, 0, l * 10 + 1);
          const uint64_t dW = make_desc_sw128(sbase + kOffW + (uint32_t)(l - 2) * kOperandBytes);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t o = (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
            umma_bf16_p(acc, dV + o, dW + o, idesc, ks > 0, leader);
            umma_bf16_p(acc + 128u, dT + o, dW + o, idesc, ks > 0, leader);
          }
          tc_commit_p(bar(kBarAccFull + s), leader);
          if (l == 4) tc_commit_p(bar(kBarSlotFree + s), leader);
          TRACE(1, s, 0, l * 10 + 2);
        }
      };
      for (;;) {
        if (lay0 == 0 && cur.valid(A) && (lay1 == 0 || slt1 != (cur.tl & 1))) {
          slt0 = cur.tl & 1, lay0 = 1;
          cur.next(A, feed);
        }
        if (lay1 == 0 && cur.valid(A) && (lay0 == 0 || slt0 != (cur.tl & 1))) {
          slt1 = cur.tl & 1, lay1 = 1;
          cur.next(A, feed);
        }
        if (lay0 == 0 && lay1 == 0) break;
        if (lay0) {
          issue(slt0, lay0);
          lay0 = (lay0 == 4) ? 0 : lay0 + 1;
        }
        if (lay1) {
          issue(slt1, lay1);
          lay1 = (lay1 == 4) ? 0 : lay1 + 1;
        }
      }
    } else if (warp == kTmaWarp) {
      // =================================== TMA issuer ===================================
      const uint32_t leader = elect_one();
      const uint8_t* Wg = reinterpret_cast<const uint8_t*>(A.W);
      mbar_arrive_expect_tx_p(bar(kBarW), 3u * kOperandBytes, leader);
#pragma unroll 1
      for (int i = 0; i < 12; ++i)
        bulk_g2s_p(sbase + kOffW + (uint32_t)i * 8192u, Wg + (size_t)A.nchunks * kOperandBytes + (size_t)i * 8192, 8192u,
                   bar(kBarW), leader);
      const int n_chunks = (d > 64) ? 4 : 2;  // K panels of the W1 image that layer 1 reads (16 KB each)
      uint32_t ph_free = 0, used = 0;
      // unit hand-out: the CTA's k-th unit comes from the global counter; it is fetched one unit ahead (the atomic's
      // round trip stays off every critical path) and published through the ring (see UnitFeed)
      volatile int* ring = reinterpret_cast<volatile int*>(smem + kOffUnits);
      auto publish = [&](uint32_t k) {
        if (leader) {
          ring[k & 3u] = atomicAdd(A.unit_counter, 1);
          mbar_arrive(bar(kBarUnit + (int)(k & 3u)));  // release: the ring store is visible to whoever sees the phase
        }
        __syncwarp();
      };
      publish(0);
      TileCursor cur;
      for (cur.init(A, feed); cur.valid(A); cur.next(A, feed)) {
        if (cur.tl == 0) publish(cur.k + 1u);  // a valid unit always announces its successor (possibly the terminator)
        const int s = cur.tl & 1;
        const uint32_t bit = 1u << s;
        const uint32_t dst = sbase + kOffSlot + (uint32_t)s * kSlotBytes + kOperandBytes;
        const int nch = MULTI ? A.nchunks : 1;
        for (int cch = 0; cch < nch; ++cch) {
          if (used & bit) {
            mbar_wait(bar(kBarSlotFree + s), (ph_free >> s) & 1u, 20);
            ph_free ^= bit;
          }
          used |= bit;
          const int np = MULTI ? 4 : n_chunks;
          mbar_arrive_expect_tx_p(bar(kBarW1Full + s), (uint32_t)np * 8192u, leader);
          for (int i = 0; i < np; ++i)
            bulk_g2s_p(dst + (uint32_t)i * 8192u, Wg + (size_t)cch * kOperandBytes + (size_t)i * 8192, 8192u,
                       bar(kBarW1Full + s), leader);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 8) {
    // =================================== epilogue warps ===================================
    // All eight warps work on ONE tile at a time: warp (q, h) owns TMEM lane quarter q (rows 32q..32q+31) and the
    // column half h of every pass, and the passes of the two slots are interleaved in the order in which the MMA
    // issuer produces their accumulators: P1(A) P1(B) P2(A) P2(B) ... P4(A) P4(B) for the tile pair (A, B) =
    // (slot 0, slot 1).  While the warps are busy with one slot the tensor pipe runs the other slot's layer, so
    // in steady state nobody waits; with four warps per slot (each slot on its own) the epilogue warps sat idle
    // 40 % of the time and a pass took twice as long.
    if constexpr (kEpiRegs > 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int q = warp & 3, h = (warp - kEpiWarp0) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tacc0 = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* const hv_row0 = smem + kOffSlot + (uint32_t)r * 128u;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    // the chunked d > 128 path is bound by its generator (issue slots), not by the XU pipe: all exponentials on MUFU there
    constexpr int kPoly = MULTI ? 0 : kEpiPolyMask;
    float2* exch = reinterpret_cast<float2*>(smem + kOffExch);  // [128] output-layer partial sums of the other half
    float* coef_s = reinterpret_cast<float*>(smem + kOffCoef);
    uint32_t ph_acc = 0, j = 0;  // ph_acc bit s = slot s
    // MULTI: Bismut sums of first-layer chunk `ew` (lane = octet x row parity), see kMultiEpiBismut
    [[maybe_unused]] float4 Ga = make_float4(0.f, 0.f, 0.f, 0.f), Gb = make_float4(0.f, 0.f, 0.f, 0.f);
    [[maybe_unused]] uint32_t ph_coef_e = 0, ph_full = 0;
    [[maybe_unused]] const int ew = warp - kEpiWarp0;
    [[maybe_unused]] auto epi_bismut = [&](uint32_t jt) {
      const uint32_t cbuf = jt % 3u, rs = jt & 3u;
      mbar_wait(bar(kBarCoef + (int)cbuf), (ph_coef_e >> cbuf) & 1u, 42);
      ph_coef_e ^= 1u << cbuf;
      mbar_wait(bar(kBarRingFull + (int)rs), (ph_full >> rs) & 1u, 43);  // (acquire: the producers' ring stores)
      ph_full ^= 1u << rs;
      const int bo = lane & 15, rh = lane >> 4, oct = 16 * ew + bo;
      if (ew < A.nchunks && 8 * oct < d) {
        const float* cf_s = coef_s + cbuf * 128u;
        const uint4* zr = A.zring + (size_t)blockIdx.x * ((size_t)4 * A.nchunks * 2048) +
                          ((size_t)rs * A.nchunks + ew) * 2048 + bo;
        constexpr int kBatch = 16;  // independent 16-byte loads in flight (L2, or DRAM when the ring was evicted)
#pragma unroll 1
        for (int i0 = 0; i0 < 64; i0 += kBatch) {
          uint4 zb[kBatch];
#pragma unroll
          for (int i = 0; i < kBatch; ++i) zb[i] = __ldcg(zr + (rh + 2 * (i0 + i)) * 16);
#pragma unroll
          for (int i = 0; i < kBatch; ++i) {
            const float cf = cf_s[rh + 2 * (i0 + i)];
            const uint4 z = zb[i];
            const float2 p0 = h2_unpack(z.x), p1 = h2_unpack(z.y), p2 = h2_unpack(z.z), p3 = h2_unpack(z.w);
            Ga.x = fmaf(cf, p0.x, Ga.x);
            Ga.y = fmaf(cf, p0.y, Ga.y);
            Ga.z = fmaf(cf, p1.x, Ga.z);
            Ga.w = fmaf(cf, p1.y, Ga.w);
            Gb.x = fmaf(cf, p2.x, Gb.x);
            Gb.y = fmaf(cf, p2.y, Gb.y);
            Gb.z = fmaf(cf, p3.x, Gb.z);
            Gb.w = fmaf(cf, p3.y, Gb.w);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarRingFree + (int)rs));  // the producers may overwrite the ring slot
    };
    TRACE_DECL(1 + h);
#ifdef DPI_TC_TRACE
    if (q != 0 || lane != 0) trace_p = nullptr;
#endif
    auto wait_acc = [&](int s) {
      mbar_wait(bar(kBarAccFull + s), (ph_acc >> s) & 1u, 30 + s);
      ph_acc ^= 1u << s;
      tc_fence_after();
    };
    auto pass_done = [&](int s) {  // next A operand written (generic -> async proxy), accumulators read
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarAReady + s));
    };
    // per-point scalars are fetched one unit ahead so that their latency is off the critical path
    float t_n = 0.f, fb_n = 0.f;
    int64_t u_n = feed.get(0);
    if (u_n < A.n_units) {
      const int64_t pn = u_n / UP;
      t_n = __ldg(A.tx + pn * (int64_t)(1 + d));
      fb_n = __ldg(A.fb + pn);
    }
    for (uint32_t k = 0;; ++k) {
      const int64_t u = u_n;
      if (u >= A.n_units) break;
      const int64_t p = u / UP;
      const int g = (int)(u % UP);
      const int tiles = min(kTcGroupTiles, A.tiles_per_point - g * kTcGroupTiles);
      const float t = t_n, fbp = fb_n;
      const float dT = horizon(P, t);
      const float* c0g = A.c0 + p * 128;  // already scaled by log2e (baseline pass)
      u_n = feed.get(k + 1u);  // published when the TMA warp started on unit k
      if (u_n < A.n_units) {
        const int64_t pn = u_n / UP;
        t_n = __ldg(A.tx + pn * (int64_t)(1 + d));
        fb_n = __ldg(A.fb + pn);
      }
      float fsum = 0.f;  // value-label partial of the tiles this warp finalises (slot h)
      for (int tl = 0; tl < tiles; tl += 2) {
        const bool hasB = tl + 1 < tiles;
        // per-row scalars of the two tiles: time sample, diffusion scale
        float a_[2], s2_[2];
        bool valid_[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int m = (g * kTcGroupTiles + tl + s) * 128 + r;
          valid_[s] = m < M_I && (s == 0 || hasB);
          float uu = 0.5f;
          if (valid_[s])
            uu = DET ? __ldg(A.u_s + p * M_I + m)
                     : u01(philox_at(A.keys, (uint64_t)(A.point_base + p), kStreamS, (uint32_t)m, 0u).x);
          a_[s] = sqrtf(uu * dT) * P.sigma;
          s2_[s] = uu * dT + t;
        }
        float pu = 0.f, pD = 0.f;
        // ---- first pass: z1' = log2e z1: the time column w_t' and c0' carry the factor already, the accumulator
        //      (W_x Z) does not
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 0 || hasB) {
            TRACE(2, s, j + s, 0);
            wait_acc(s);
            TRACE(2, s, j + s, 1);
            uint8_t* const hv = hv_row0 + (uint32_t)s * kSlotBytes;
            epi_pass<kPassFirst, kAcc16First && !MULTI, kPoly>(C, tacc0 + (uint32_t)s * 256u, hv, hv + kOperandBytes, rx, h, c0g, 0, s2_[s],
                                 a_[s] * kLog2e, pu, pD);
            pass_done(s);
            TRACE(2, s, j + s, 2);
          }
        }
        // ---- hidden passes
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (s == 0 || hasB) {
              wait_acc(s);
              TRACE(2, s, j + s, (l + 1) * 10 + 1);
              uint8_t* const hv = hv_row0 + (uint32_t)s * kSlotBytes;
              epi_pass<kPassHidden, kAcc16, kPoly>(C, tacc0 + (uint32_t)s * 256u, hv, hv + kOperandBytes, rx, h, c0g, l, 0.f, 0.f, pu,
                                    pD);
              pass_done(s);
              TRACE(2, s, j + s, (l + 1) * 10 + 2);
            }
          }
        }
        // ---- output pass: each half takes the dot products over its 64 columns; the warp with h == slot adds the
        //      other half's partials (one exchange buffer, the direction alternates with the slot) and finishes
        //      the row: output layer, f (equations.py:297-302), Bismut coefficient
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 0 || hasB) {
            wait_acc(s);
            TRACE(2, s, j + s, 31);
            pu = pD = 0.f;
            uint8_t* const hv = hv_row0 + (uint32_t)s * kSlotBytes;
            epi_pass<kPassOut, kAcc16, kPoly>(C, tacc0 + (uint32_t)s * 256u, hv, hv + kOperandBytes, rx, h, c0g, 2, 0.f, 0.f, pu, pD);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarAReady + s));  // accumulators of the slot are free for the next tile
            if (h != s) exch[r] = make_float2(pu, pD);
            named_bar_sync(1 + q, 64);  // the two warps of lane quarter q
            if (h == s) {
              const float2 o = exch[r];
              const float uval = (pu + o.x) + C.bout, Dv = pD + o.y;
              const float f = P.sigma * (P.cha_k * uval - P.cha_c) * (P.sigma * Dv);
#ifdef DPI_TC_DEBUG_UD
              if (A.trace && valid_[s])
                reinterpret_cast<float2*>(A.trace)[p * M_I + (g * kTcGroupTiles + tl + s) * 128 + r] =
                    make_float2(uval, Dv);
#endif
              if (valid_[s]) fsum += dT * f;
              const uint32_t cb = (j + (uint32_t)s) % 3u;
              coef_s[cb * 128u + r] = valid_[s] ? dT * (f - fbp) / a_[s] : 0.f;
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(kBarCoef + cb));  // (release: orders the warp's coefficient stores)
            }
            TRACE(2, s, j + s, 32);
          }
        }
        if constexpr (MULTI && kMultiEpiBismut) {
          epi_bismut(j);
          if (hasB) epi_bismut(j + 1u);
        }
        j += hasB ? 2u : 1u;
      }
      if constexpr (MULTI && kMultiEpiBismut) {  // unit partial: [chunk][row parity][128]
        if (ew < A.nchunks) {
          float* gp_ = A.partG + u * kTcPartGFloatsPerUnit + (int64_t)(ew * 2 + (lane >> 4)) * 128 + 8 * (lane & 15);
          *reinterpret_cast<float4*>(gp_) = Ga;
          *reinterpret_cast<float4*>(gp_ + 4) = Gb;
        }
        Ga = Gb = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float ws = warp_sum(fsum);
      if (lane == 0) A.partF[u * kTcPartFFloatsPerUnit + h * 4 + q] = ws;
    }
  } else {
    // =================================== increment producers ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kProdRegs));
    if constexpr (MULTI) {
      // ---- 128 < d <= 1024: chunked first layer, increments kept in a global ring (see the kernel comment) ----
      const int tl_id = threadIdx.x - 32 * kProdWarp0;
      const int o = tl_id & 15, c = tl_id >> 4;  // octet inside the 128-dim chunk, row class: rows c + 16 i
      const int pw = tl_id >> 5;                 // producer warp 0..7 = the chunk whose Bismut sums it owns
      const int nch = A.nchunks;
      const bool vec_ok = (d & 3) == 0;
      uint32_t zoff[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t rr = (uint32_t)(c + 16 * i);
        zoff[i] = (uint32_t)(o >> 3) * kPanelBytes + rr * 128u + (((uint32_t)(o & 7) ^ (rr & 7u)) << 4);
      }
      const float* coef_s = reinterpret_cast<const float*>(smem + kOffCoef);
      float4 Ga = make_float4(0.f, 0.f, 0.f, 0.f), Gb = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t ph_free = 0, used = 0, ph_coef = 0, j = 0;
      [[maybe_unused]] uint32_t ph_ring = 0;
      int pending = 0;
      int64_t u1 = 0, u2 = 0;  // units of the pending tiles (1 = newer, 2 = older)
      int m1 = 0, m2 = 0;      // their first samples
      bool last1 = false, last2 = false;
      uint32_t j1 = 0, j2 = 0;
      // fp16-rounded increments of dims [8 oct, 8 oct + 8) of sample m of point pp: what the MLP saw
      auto draw8 = [&](int64_t pp, int m, int oct, bool ok) -> uint4 {
        float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
        if (ok) {
          if (DET) {
            const float* row = A.zS + (pp * M_I + m) * d;
            za = load_row4(row, 8 * oct, d, vec_ok);
            zb = load_row4(row, 8 * oct + 4, d, vec_ok);
          } else {
#ifdef DPI_DBG_NO_RNG  // sensitivity builds only (wrong numbers)
            const uint32_t hsh = (uint32_t)m * 2654435761u + (uint32_t)oct * 40503u;
            za = make_float4(__uint_as_float(0x3f000000u | (hsh & 0x7fffffu)) - 0.75f,
                             __uint_as_float(0x3f000000u | ((hsh >> 3) & 0x7fffffu)) - 0.75f,
                             __uint_as_float(0x3f000000u | ((hsh >> 6) & 0x7fffffu)) - 0.75f,
                             __uint_as_float(0x3f000000u | ((hsh >> 9) & 0x7fffffu)) - 0.75f);
            zb = za;
#else
            normal8<true>(A.keys, (uint64_t)(A.point_base + pp), kStreamIntegral, (uint32_t)m, (uint32_t)oct, za, zb);
#endif
          }
        }
        return ok ? make_uint4(pack_f16x2(za.x, za.y), pack_f16x2(za.z, za.w), pack_f16x2(zb.x, zb.y), pack_f16x2(zb.z, zb.w))
                  : make_uint4(0u, 0u, 0u, 0u);
      };
      // The fp16 increments of the tiles whose coefficients are outstanding wait in a per-CTA ring in global memory
      // (4 tiles x chunks x 32 KB, L2 resident for the most part): written once by the generating threads, read
      // back by warp pw for chunk pw (lane = octet x row parity, 64 rows each).  Re-drawing them from the Philox
      // counters instead cost as much as the first draw -- and the generator is what bounds this kernel.
      // Four slots although only three tiles are ever live: a warp that is done with the sums of tile j-2 starts
      // writing tile j+1 while a slower warp may still be reading tile j-2 (three slots would alias them); the
      // per-chunk increments barrier keeps the warps within one tile of each other.
      uint4* const zring = A.zring + (size_t)blockIdx.x * ((size_t)4 * nch * 2048);
      auto bismut = [&](uint32_t jt, int64_t ut, int mt0, bool last) {
        const uint32_t cbuf = jt % 3u;
        mbar_wait(bar(kBarCoef + (int)cbuf), (ph_coef >> cbuf) & 1u, 40);
        ph_coef ^= 1u << cbuf;
        const int bo = lane & 15, rh = lane >> 4, oct = 16 * pw + bo;
        if (pw < nch && 8 * oct < d) {
          const float* cf_s = coef_s + cbuf * 128u;
          const uint4* zr = zring + ((size_t)(jt & 3u) * nch + pw) * 2048 + bo;
          // The ring is larger than L2 (4 tiles x nch x 32 KB per CTA), so a read-back is a DRAM round trip: kBisBatch
          // independent 16-byte loads are issued before the first is consumed (4 in flight made the read-back half of
          // the producers' time in ncu's stall samples), on top of the L2 prefetch issued a tile earlier.
          constexpr int kBisBatch = 16;
          for (int i0 = 0; i0 < 64; i0 += kBisBatch) {
            uint4 zb[kBisBatch];
#pragma unroll
            for (int i = 0; i < kBisBatch; ++i)
#ifdef DPI_DBG_MULTI_NOREAD  // sensitivity builds only (wrong numbers)
              zb[i] = make_uint4(i0, i, rh, bo);
#else
              zb[i] = __ldcg(zr + (rh + 2 * (i0 + i)) * 16);
#endif
#pragma unroll
            for (int i = 0; i < kBisBatch; ++i) {
              const float cf = cf_s[rh + 2 * (i0 + i)];
              const uint4 z = zb[i];
              const float2 p0 = h2_unpack(z.x), p1 = h2_unpack(z.y), p2 = h2_unpack(z.z), p3 = h2_unpack(z.w);
              Ga.x = fmaf(cf, p0.x, Ga.x);
              Ga.y = fmaf(cf, p0.y, Ga.y);
              Ga.z = fmaf(cf, p1.x, Ga.z);
              Ga.w = fmaf(cf, p1.y, Ga.w);
              Gb.x = fmaf(cf, p2.x, Gb.x);
              Gb.y = fmaf(cf, p2.y, Gb.y);
              Gb.z = fmaf(cf, p3.x, Gb.z);
              Gb.w = fmaf(cf, p3.y, Gb.w);
            }
          }
        }
        (void)mt0;
        if (last) {  // unit partial: [chunk][row parity][128]
          if (pw < nch) {
            float* gp_ = A.partG + ut * kTcPartGFloatsPerUnit + (int64_t)(pw * 2 + rh) * 128 + 8 * bo;
            *reinterpret_cast<float4*>(gp_) = Ga;
            *reinterpret_cast<float4*>(gp_ + 4) = Gb;
          }
          Ga = Gb = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      for (uint32_t k = 0;; ++k) {
        const int64_t u = feed.get(k);
        if (u >= A.n_units) break;
        const int64_t p = u / UP;
        const int g = (int)(u % UP);
        const int tiles = min(kTcGroupTiles, A.tiles_per_point - g * kTcGroupTiles);
        for (int tl = 0; tl < tiles; ++tl, ++j) {
          const int m0 = (g * kTcGroupTiles + tl) * 128;
          const int nvalid = min(128, M_I - m0);
          const int s = tl & 1;
          const uint32_t bit = 1u << s;
          uint8_t* const zdst = smem + kOffSlot + (uint32_t)s * kSlotBytes;
          if constexpr (kMultiEpiBismut) {
            if (j >= 4u) {  // the epilogue warps are done with the tile that used this ring slot
              const uint32_t rs = j & 3u;
              mbar_wait(bar(kBarRingFree + (int)rs), (ph_ring >> rs) & 1u, 44);
              ph_ring ^= 1u << rs;
            }
          }
#ifndef DPI_MULTI_NO_PREFETCH
          if (!kMultiEpiBismut && pending == 2 && pw < nch) {
            // the chunk this warp sums once this tile is generated (32 KB, contiguous): pull it towards L2 now
            const char* pf = reinterpret_cast<const char*>(zring + ((size_t)(j2 & 3u) * nch + pw) * 2048);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)(lane + 32 * i) * 128));
          }
#endif
          for (int cch = 0; cch < nch; ++cch) {
            const int oct = 16 * cch + o;
            const bool col_ok = 8 * oct < d;
            uint4 zc[8];
#if !defined(DPI_NO_PHILOX_X2) && !defined(DPI_DBG_NO_RNG)
            if (!DET && col_ok && nvalid == 128) {  // full tile: two rows per step, Philox rounds interleaved
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                float4 za0, zb0, za1, zb1;
                normal8x2<true>(A.keys, (uint64_t)(A.point_base + p), kStreamIntegral, (uint32_t)(m0 + c + 16 * i),
                                (uint32_t)(m0 + c + 16 * (i + 1)), (uint32_t)oct, za0, zb0, za1, zb1);
                zc[i] = make_uint4(pack_f16x2(za0.x, za0.y), pack_f16x2(za0.z, za0.w), pack_f16x2(zb0.x, zb0.y),
                                   pack_f16x2(zb0.z, zb0.w));
                zc[i + 1] = make_uint4(pack_f16x2(za1.x, za1.y), pack_f16x2(za1.z, za1.w), pack_f16x2(zb1.x, zb1.y),
                                       pack_f16x2(zb1.z, zb1.w));
              }
            } else
#endif
            {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = c + 16 * i;
                zc[i] = draw8(p, m0 + rr, oct, col_ok && rr < nvalid);
              }
            }
            if (used & bit) {
              mbar_wait(bar(kBarSlotFree + s), (ph_free >> s) & 1u, 41);  // previous chunk / tile has left hV / hT
              ph_free ^= bit;
            }
            used |= bit;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(zdst + zoff[i]) = zc[i];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarZFull + s));
            {  // keep the chunk for the Bismut sums (rows c + 16 i, octet o of ring slot j % 4)
              uint4* zw = zring + ((size_t)(j & 3u) * nch + cch) * 2048 + o;
#pragma unroll
#ifndef DPI_DBG_MULTI_NOSTORE  // sensitivity builds only (wrong numbers)
              for (int i = 0; i < 8; ++i) __stcg(zw + (c + 16 * i) * 16, zc[i]);
#endif
            }
          }
          __threadfence_block();
          if constexpr (kMultiEpiBismut) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarRingFull + (int)(j & 3u)));  // (release: the warp's ring stores of tile j)
          } else {
            if (pending == 2) bismut(j2, u2, m2, last2);
            j2 = j1, u2 = u1, m2 = m1, last2 = last1;
            j1 = j, u1 = u, m1 = m0, last1 = (tl == tiles - 1);
            if (pending < 2) ++pending;
          }
        }
      }
      if constexpr (!kMultiEpiBismut) {
        if (pending == 2) bismut(j2, u2, m2, last2);
        if (pending >= 1) bismut(j1, u1, m1, last1);
      }
    } else {
    // All 8 warps draw every tile: thread (o, c) owns the 8 dims [8o, 8o+8) -- one Philox block, one 16-byte operand
    // chunk -- of the rows r = c + cls i.  A tile's increments are generated into registers (no dependence on shared
    // memory), stored into the slot's hV buffer once the slot's previous tile has left it, and kept until the tile's
    // Bismut coefficients arrive two tiles later.
    const int tl_id = threadIdx.x - 32 * kProdWarp0;
    const int no = A.no, cls = A.cls;
    const int o = tl_id % no, c = tl_id / no;
    const bool active = c < cls;
    const bool vec_ok = (d & 3) == 0;
    uint32_t zoff[ITEMS];  // byte offset of item i in a slot's hV image (K panel, row, swizzled 16-byte chunk)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t rr = (uint32_t)(c + cls * i);
      zoff[i] = (uint32_t)(o >> 3) * kPanelBytes + rr * 128u + (((uint32_t)(o & 7) ^ (rr & 7u)) << 4);
    }
    const float* coef_s = reinterpret_cast<const float*>(smem + kOffCoef);
    uint4 zc[ITEMS], z1[ITEMS], z2[ITEMS];  // fp16x8 per item: current tile and the two tiles in flight
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) z1[i] = z2[i] = make_uint4(0u, 0u, 0u, 0u);
    float4 Ga = make_float4(0.f, 0.f, 0.f, 0.f), Gb = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t ph_free = 0, used = 0, ph_coef = 0, j = 0;
    int pending = 0;
    int64_t u1 = 0, u2 = 0;      // units of the pending tiles (1 = newer, 2 = older)
    bool last1 = false, last2 = false;
    uint32_t j1 = 0, j2 = 0;
    TRACE_DECL(3);
#ifdef DPI_TC_TRACE
    if (tl_id != 0) trace_p = nullptr;
#endif
    // Bismut accumulation of a finished tile: G_j += coef_m Z_mj (fp16 Z, exactly what the MLP saw)
    auto bismut = [&](const uint4 (&z)[ITEMS], uint32_t jt, int64_t ut, bool last) {
      const uint32_t cbuf = jt % 3u;
      mbar_wait(bar(kBarCoef + (int)cbuf), (ph_coef >> cbuf) & 1u, 40);
      ph_coef ^= 1u << cbuf;
      if (active) {
        const float* cf_s = coef_s + cbuf * 128u;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          const int rr = c + cls * i;
          if (rr < 128) {
            const float cf = cf_s[rr];
            const float2 p0 = h2_unpack(z[i].x), p1 = h2_unpack(z[i].y), p2 = h2_unpack(z[i].z), p3 = h2_unpack(z[i].w);
            Ga.x = fmaf(cf, p0.x, Ga.x);
            Ga.y = fmaf(cf, p0.y, Ga.y);
            Ga.z = fmaf(cf, p1.x, Ga.z);
            Ga.w = fmaf(cf, p1.y, Ga.w);
            Gb.x = fmaf(cf, p2.x, Gb.x);
            Gb.y = fmaf(cf, p2.y, Gb.y);
            Gb.z = fmaf(cf, p3.x, Gb.z);
            Gb.w = fmaf(cf, p3.y, Gb.w);
          }
        }
        if (last) {
          float* gp_ = A.partG + (ut * kTcMaxClasses + c) * 128 + 8 * o;
          *reinterpret_cast<float4*>(gp_) = Ga;
          *reinterpret_cast<float4*>(gp_ + 4) = Gb;
          Ga = Gb = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    for (uint32_t k = 0;; ++k) {
      const int64_t u = feed.get(k);
      if (u >= A.n_units) break;
      const int64_t p = u / UP;
      const int g = (int)(u % UP);
      const int tiles = min(kTcGroupTiles, A.tiles_per_point - g * kTcGroupTiles);
      const uint64_t gp = (uint64_t)(A.point_base + p);
      for (int tl = 0; tl < tiles; ++tl, ++j) {
        const int m0 = (g * kTcGroupTiles + tl) * 128;
        const int nvalid = min(128, M_I - m0);
        TRACE(3, tl & 1, j, 0);
        // (measured: -3.7 % on the d > 128 kernel, whose producers are latency bound; +0.5 % here, where the epilogue
        // warps want the same issue slots -- so this path keeps one block at a time unless DPI_PHILOX_X2_ALL is set)
#if defined(DPI_PHILOX_X2_ALL) && !defined(DPI_DBG_NO_RNG)
        constexpr bool kX2 = !DET;
#else
        constexpr bool kX2 = false;
#endif
        if (kX2 && active && nvalid == 128) {  // full tile: two rows per step, Philox rounds interleaved (normal8x2)
          auto pk = [](const float4& za, const float4& zb) {
            return make_uint4(pack_f16x2(za.x, za.y), pack_f16x2(za.z, za.w), pack_f16x2(zb.x, zb.y), pack_f16x2(zb.z, zb.w));
          };
#pragma unroll
          for (int i = 0; i + 1 < ITEMS; i += 2) {
            float4 za0, zb0, za1, zb1;
            normal8x2(A.keys, gp, kStreamIntegral, (uint32_t)(m0 + c + cls * i), (uint32_t)(m0 + c + cls * (i + 1)),
                      (uint32_t)o, za0, zb0, za1, zb1);
            zc[i] = pk(za0, zb0);  // a row >= 128 of the last items is drawn like any other and never stored
            zc[i + 1] = pk(za1, zb1);
          }
          if constexpr (ITEMS & 1) {
            float4 za, zb;
            normal8(A.keys, gp, kStreamIntegral, (uint32_t)(m0 + c + cls * (ITEMS - 1)), (uint32_t)o, za, zb);
            zc[ITEMS - 1] = pk(za, zb);
          }
        } else if (active) {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            const int rr = c + cls * i;
            float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
            if (DET) {
              if (rr < nvalid) {
                const float* row = A.zS + ((int64_t)p * M_I + m0 + rr) * d;
                za = load_row4(row, 8 * o, d, vec_ok);
                zb = load_row4(row, 8 * o + 4, d, vec_ok);
              }
            } else {
#ifdef DPI_DBG_NO_RNG
              {
                const uint32_t hsh = (uint32_t)(m0 + rr) * 2654435761u + (uint32_t)o * 40503u;
                za = make_float4(__uint_as_float(0x3f000000u | (hsh & 0x7fffffu)) - 0.75f,
                                 __uint_as_float(0x3f000000u | ((hsh >> 3) & 0x7fffffu)) - 0.75f,
                                 __uint_as_float(0x3f000000u | ((hsh >> 6) & 0x7fffffu)) - 0.75f,
                                 __uint_as_float(0x3f000000u | ((hsh >> 9) & 0x7fffffu)) - 0.75f);
                zb = za;
              }
#else
              normal8(A.keys, gp, kStreamIntegral, (uint32_t)(m0 + rr), (uint32_t)o, za, zb);
#endif
              // dims >= d of the last octet stay as drawn: their W1 columns are zero and their Bismut sums are never
              // read, so masking them would only cost instructions
            }
            const bool ok = rr < nvalid;
            zc[i] = ok ? make_uint4(pack_f16x2(za.x, za.y), pack_f16x2(za.z, za.w), pack_f16x2(zb.x, zb.y),
                                    pack_f16x2(zb.z, zb.w))
                       : make_uint4(0u, 0u, 0u, 0u);
          }
        }
        const int s = tl & 1;
        const uint32_t bit = 1u << s;
        TRACE(3, s, j, 1);
        if (used & bit) {
          mbar_wait(bar(kBarSlotFree + s), (ph_free >> s) & 1u, 41);  // the slot's previous tile has left hV / hT
          ph_free ^= bit;
        }
        used |= bit;
        TRACE(3, s, j, 2);
        if (active) {
          uint8_t* const zdst = smem + kOffSlot + (uint32_t)s * kSlotBytes;
#pragma unroll
          for (int i = 0; i < ITEMS; ++i) {
            if (c + cls * i < 128) *reinterpret_cast<uint4*>(zdst + zoff[i]) = zc[i];
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarZFull + s));
        TRACE(3, s, j, 3);
        if (pending == 2) bismut(z2, j2, u2, last2);
        TRACE(3, s, j, 4);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
          z2[i] = z1[i];
          z1[i] = zc[i];
        }
        j2 = j1, u2 = u1, last2 = last1;
        j1 = j, u1 = u, last1 = (tl == tiles - 1);
        if (pending < 2) ++pending;
      }
    }
    if (pending == 2) bismut(z2, j2, u2, last2);
    if (pending >= 1) bismut(z1, j1, u1, last1);
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef DPI_TC_TRACE
  if (A.trace && threadIdx.x == 0) {  // per-CTA wall time (globaltimer, ns) and SM id
    unsigned long long t1;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long* q = A.trace + 5 * 2 * kTraceCap + 4 * blockIdx.x;
    q[0] = (long long)cta_t0, q[1] = (long long)t1, q[2] = smid, q[3] = 1;
  }
#endif
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// y[p] += (sum of the unit partials in a fixed order) / M_I, then clip (data.py:222)
__global__ void __launch_bounds__(128)
cha_tc_finish_kernel(const float* __restrict__ partG, const float* __restrict__ partF, int units_per_point, int cls, int d,
                     int M_I, float bound, float* __restrict__ y, int* __restrict__ nonfinite) {
  const int64_t p = blockIdx.x;
  const int j = threadIdx.x;
  const float invM = 1.f / (float)M_I;
  const bool clip = bound > 0.f && isfinite(bound);
  float* yp = y + p * (int64_t)(1 + d);
  bool bad = false;
  if (j < d) {
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int c = 0; c < cls; ++c) s += partG[((p * units_per_point + g) * kTcMaxClasses + c) * 128 + j];
    float v = yp[1 + j] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[1 + j] = v;
    bad |= !isfinite(v);
  }
  if (j == 0) {
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int qq = 0; qq < (int)kTcPartFFloatsPerUnit; ++qq)
        s += partF[(p * units_per_point + g) * kTcPartFFloatsPerUnit + qq];
    float v = yp[0] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[0] = v;
    bad |= !isfinite(v);
  }
  // an overflowing 16-bit operand (inf -> NaN through the MMA) ends up here: report it instead of returning NaN labels
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(nonfinite, 1);
}

// the same for the chunked layout (d > 128): unit partials are [chunk][row parity][128]
__global__ void __launch_bounds__(128)
cha_tc_finish_multi_kernel(const float* __restrict__ partG, const float* __restrict__ partF, int units_per_point, int d,
                           int M_I, float bound, float* __restrict__ y, int* __restrict__ nonfinite) {
  const int64_t p = blockIdx.x;
  const float invM = 1.f / (float)M_I;
  const bool clip = bound > 0.f && isfinite(bound);
  float* yp = y + p * (int64_t)(1 + d);
  bool bad = false;
  for (int j = threadIdx.x; j < d; j += 128) {
    const int cch = j >> 7, jj = j & 127;
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int rh = 0; rh < 2; ++rh)
        s += partG[(p * units_per_point + g) * kTcPartGFloatsPerUnit + (int64_t)(cch * 2 + rh) * 128 + jj];
    float v = yp[1 + j] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[1 + j] = v;
    bad |= !isfinite(v);
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int g = 0; g < units_per_point; ++g)
      for (int qq = 0; qq < (int)kTcPartFFloatsPerUnit; ++qq)
        s += partF[(p * units_per_point + g) * kTcPartFFloatsPerUnit + qq];
    float v = yp[0] + s * invM;
    if (clip) v = fminf(fmaxf(v, -bound), bound);
    yp[0] = v;
    bad |= !isfinite(v);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(nonfinite, 1);
}

// dst = fp16 SWIZZLE_128B K-major image of W[n][col_off + k] (n < n_valid, k < k_valid; zero elsewhere)
__global__ void pack_sw128_kernel(const float* __restrict__ W, int ld, int col_off, int n_valid, int k_valid,
                                  uint8_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 128) return;
  const int n = i >> 7, k = i & 127;
  const float v = (n < n_valid && k < k_valid) ? W[(size_t)n * ld + col_off + k] : 0.f;
  *reinterpret_cast<__half*>(dst + sw128_offset((uint32_t)n, (uint32_t)k)) = __float2half_rn(v);
}

// ---- UMMA self-test: D[128x128] = A[128x128] . B[128x128]^T through the same descriptors / layouts -------------
// mode 0: fp32 accumulators; mode 1: fp16 accumulators (instruction descriptor c_format = 0) read back as raw 32-bit
// TMEM cells; mode 2: fp16 accumulators read back with tcgen05.ld ... .pack::16b (two adjacent columns per register)
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const uint8_t* __restrict__ Aimg,
                                                               const uint8_t* __restrict__ Bimg, float* __restrict__ D,
                                                               int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t barW = sbase + 2 * kOperandBytes, barD = barW + 8, tptr = barW + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (sbase & 1023u) __trap();
    mbar_init(barW, 1);
    mbar_init(barD, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 2 * kOperandBytes + 16);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(barW, 2u * kOperandBytes);
    bulk_g2s(sbase, Aimg, kOperandBytes, barW);
    bulk_g2s(sbase + kOperandBytes, Bimg, kOperandBytes, barW);
    mbar_wait(barW, 0, 100);
    tc_fence_after();
    const uint32_t idesc = mode == 0 ? make_idesc_f16(128, 128) : (make_idesc_f16(128, 128) & ~(3u << 4));
    for (int ks = 0; ks < 8; ++ks)
      umma_bf16(tmem, desc_kstep(sbase, ks), desc_kstep(sbase + kOperandBytes, ks), idesc, ks > 0);
    tc_commit(barD);
  }
  __syncwarp();
  mbar_wait(barD, 0, 101);
  tc_fence_after();
  const int r = warp * 32 + lane;
  if (mode == 2) {
    for (int col0 = 0; col0 < 128; col0 += 32) {
      uint32_t v[16];
      tmem_ld16_pack(tmem + ((uint32_t)(warp * 32) << 16) + col0, v);  // 32 columns -> 16 registers of two halves
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) {
        const float2 f = h2_unpack(v[i]);
        D[r * 128 + col0 + 2 * i] = f.x;
        D[r * 128 + col0 + 2 * i + 1] = f.y;
      }
    }
  } else {
    for (int col0 = 0; col0 < 128; col0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) D[r * 128 + col0 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

template <int ITEMS, bool DET, bool MULTI = false>
int launch_tc_t(const ProblemDev& P, const TcNetDev& N, const TcArgs& A, int grid, cudaStream_t st) {
  auto k = cha_tc_kernel<ITEMS, DET, MULTI>;
  DPI_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  k<<<grid, kTcThreads, kTcSmemBytes, st>>>(P, N.consts, A);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

}  // namespace

int tc_pack_network(const dpi_network& net, const ProblemDev& P, const MlpDev& M, std::vector<void*>& owner,
                    TcNetDev* out, bool* ready, int64_t* launches, cudaStream_t st) {
  *ready = false;
  const int L = net.n_main;
  // this kernel: Linear(1+d,H1), Linear(H1,H2), Linear(H2,H3), Linear(H3,H4), Linear(H4,1) with every H_l <= 128.
  // Narrower layers are zero-padded to the 128 x 128 operand images: a padded unit has zero weights and bias, so its
  // activation, ELU' * tangent and output weight are all exactly zero.  (M.HP == 128 is what the baseline pass and the
  // per-point c0 rows are laid out for; dpi_set_network picks it for every five-layer Cha network.)
  if (L != 5 || P.nx > 1024 || M.HP != 128) return DPI_OK;
  const int nch = (P.nx + 127) / 128;        // first-layer K chunks of 128 input dims
  for (int l = 0; l < L - 1; ++l)
    if (net.main[l].n_out > 128) return DPI_OK;
  void* img = nullptr;
  DPI_CUDA_CHECK(cudaMalloc(&img, (size_t)(nch + 3) * tc::kOperandBytes));
  owner.push_back(img);
  uint8_t* base = reinterpret_cast<uint8_t*>(img);
  for (int cch = 0; cch < nch; ++cch)  // W1 x-columns [128 cch, 128 cch + 128)
    pack_sw128_kernel<<<64, 256, 0, st>>>(net.main[0].W, 1 + P.nx, 1 + 128 * cch, net.main[0].n_out,
                                          std::min(128, P.nx - 128 * cch), base + (size_t)cch * tc::kOperandBytes);
  for (int l = 1; l <= 3; ++l)
    pack_sw128_kernel<<<64, 256, 0, st>>>(net.main[l].W, net.main[l].n_in, 0, net.main[l].n_out, net.main[l].n_in,
                                          base + (size_t)(nch - 1 + l) * tc::kOperandBytes);
  DPI_CUDA_CHECK(cudaGetLastError());
  *launches += nch + 3;
  // per-column fp32 constants travel as kernel parameters (constant bank): copy the packed vectors back once
  DPI_CUDA_CHECK(cudaStreamSynchronize(st));
  TcNetDev T{};
  // activations are carried scaled by log2(e) (see epi_pass): scale the vectors that add to / read from them
  std::vector<float> wt(128), wsum(128), bh(128), wout(128);
  DPI_CUDA_CHECK(cudaMemcpy(wt.data(), M.wt, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(wsum.data(), M.wsum, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(wout.data(), M.wout, 128 * sizeof(float), cudaMemcpyDeviceToHost));
  DPI_CUDA_CHECK(cudaMemcpy(&T.consts.bout, M.bout, sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 128; ++i) T.consts.wt[i] = wt[i] * kLog2e;
  for (int i = 0; i < 64; ++i) {
    T.consts.wsum16[i] = h2_const(wsum[2 * i] * kLog2e, wsum[2 * i + 1] * kLog2e);
    T.consts.wt16[i] = h2_const(wt[2 * i] * kLog2e, wt[2 * i + 1] * kLog2e);
    T.consts.wout16[i] = h2_const(wout[2 * i] * 0.6931471805599453f, wout[2 * i + 1] * 0.6931471805599453f);
  }
  for (int l = 0; l < 3; ++l) {
    DPI_CUDA_CHECK(cudaMemcpy(bh.data(), M.bh[l], 128 * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 64; ++i) T.consts.b16[l][i] = h2_const(bh[2 * i] * kLog2e, bh[2 * i + 1] * kLog2e);
  }
  T.W = img;
  T.nx = P.nx;
  T.nchunks = nch;
  T.ksteps1 = nch > 1 ? 8 : (P.nx + 15) / 16;
  T.no = nch > 1 ? 16 : (P.nx + 7) / 8;
  T.cls = 256 / T.no < kTcMaxClasses ? 256 / T.no : kTcMaxClasses;  // row classes of the 256 producer threads
  int dev = 0, sms = 0;
  DPI_CUDA_CHECK(cudaGetDevice(&dev));
  DPI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  T.num_sms = sms;
  *out = T;
  *ready = true;
  return DPI_OK;
}

int64_t tc_units_per_point(int M_I) {
  const int tiles = (M_I + 127) / 128;
  return (tiles + kTcGroupTiles - 1) / kTcGroupTiles;
}

int launch_cha_tc_path(const ProblemDev& P, const TcNetDev& N, const float* tx, int64_t n, int M_I, uint64_t seed,
                       int64_t base, const float* u_s, const float* zS, const float* fb, const float* c0, float bound,
                       float* partG, float* partF, int* unit_counter, void* zring, int* nonfinite, float* y,
                       cudaStream_t st) {
  if (n <= 0) return DPI_OK;
  if (n * tc_units_per_point(M_I) > 0x7fffff00LL) {
    set_error("tensor-core path: %lld points x %lld units exceed one launch; shard the call", (long long)n,
              (long long)tc_units_per_point(M_I));
    return DPI_ERR_INVALID;
  }
  TcArgs A{};
  A.unit_counter = unit_counter;
  A.zring = reinterpret_cast<uint4*>(zring);
  if (N.nchunks > 1 && !zring) return DPI_ERR_INVALID;
  DPI_CUDA_CHECK(cudaMemsetAsync(unit_counter, 0, sizeof(int), st));
  A.W = N.W;
  A.tx = tx;
  A.tiles_per_point = (M_I + 127) / 128;
  A.units_per_point = (int)tc_units_per_point(M_I);
  A.n_units = n * A.units_per_point;
  A.point_base = base;
  A.seed = seed;
  A.keys = make_philox_keys(seed);
  A.u_s = u_s;
  A.zS = zS;
  A.fb = fb;
  A.c0 = c0;
  A.partG = partG;
  A.partF = partF;
  A.M_I = M_I;
  A.ksteps1 = N.ksteps1;
  A.no = N.no;
  A.cls = N.cls;
  A.nchunks = N.nchunks;
  const int grid = (int)(A.n_units < N.num_sms ? A.n_units : N.num_sms);
  int rc;
#ifdef DPI_TC_DEBUG_UD
  float2* ud_dev = nullptr;
  if (getenv("DPI_TC_UD_FILE")) {
    cudaMalloc((void**)&ud_dev, sizeof(float2) * (size_t)n * M_I);
    cudaMemsetAsync(ud_dev, 0, sizeof(float2) * (size_t)n * M_I, st);
    A.trace = reinterpret_cast<long long*>(ud_dev);
  }
#endif
#ifdef DPI_TC_TRACE
  long long* trace_dev = nullptr;
  const size_t trace_bytes = sizeof(long long) * (6 * 2 * 8192);
  if (getenv("DPI_TC_TRACE_FILE")) {
    cudaMalloc((void**)&trace_dev, trace_bytes);
    cudaMemsetAsync(trace_dev, 0, trace_bytes, st);
    A.trace = trace_dev;
  }
#endif
  const bool det = zS != nullptr;
  const int items = (128 + N.cls - 1) / N.cls;  // rows per producer thread: 4 (d <= 64) .. 8 (d > 104)
  if (N.nchunks > 1) {
    rc = det ? launch_tc_t<8, true, true>(P, N, A, grid, st) : launch_tc_t<8, false, true>(P, N, A, grid, st);
  } else
  switch (items) {
    case 4: rc = det ? launch_tc_t<4, true>(P, N, A, grid, st) : launch_tc_t<4, false>(P, N, A, grid, st); break;
    case 5: rc = det ? launch_tc_t<5, true>(P, N, A, grid, st) : launch_tc_t<5, false>(P, N, A, grid, st); break;
    case 6: rc = det ? launch_tc_t<6, true>(P, N, A, grid, st) : launch_tc_t<6, false>(P, N, A, grid, st); break;
    case 7: rc = det ? launch_tc_t<7, true>(P, N, A, grid, st) : launch_tc_t<7, false>(P, N, A, grid, st); break;
    default: rc = det ? launch_tc_t<8, true>(P, N, A, grid, st) : launch_tc_t<8, false>(P, N, A, grid, st); break;
  }
  if (rc) return rc;
#ifdef DPI_TC_TRACE
  if (trace_dev) {
    std::vector<long long> hbuf(6 * 2 * 8192);
    cudaStreamSynchronize(st);
    cudaMemcpy(hbuf.data(), trace_dev, trace_bytes, cudaMemcpyDeviceToHost);
    cudaFree(trace_dev);
    if (FILE* f = fopen(getenv("DPI_TC_TRACE_FILE"), "w")) {
      for (int r = 0; r < 5; ++r)
        for (int i = 0; i < 8192; ++i) {
          const long long code = hbuf[(size_t)r * 2 * 8192 + 2 * i], clk = hbuf[(size_t)r * 2 * 8192 + 2 * i + 1];
          if (clk) fprintf(f, "%lld %lld\n", (long long)((r == 0 || r == 4) ? 1 : (r == 3 ? 3 : 2)) * 1000000 + code, clk);
        }
      fclose(f);
      if (FILE* g = fopen((std::string(getenv("DPI_TC_TRACE_FILE")) + ".cta").c_str(), "w")) {
        const long long* q = hbuf.data() + (size_t)5 * 2 * 8192;
        for (int b = 0; b < 2048 && q[4 * b + 3]; ++b) fprintf(g, "%d %lld %lld %lld\n", b, q[4 * b], q[4 * b + 1], q[4 * b + 2]);
        fclose(g);
      }
    }
  }
#endif
#ifdef DPI_TC_DEBUG_UD
  if (ud_dev) {
    std::vector<float2> hb((size_t)n * M_I);
    cudaStreamSynchronize(st);
    cudaMemcpy(hb.data(), ud_dev, sizeof(float2) * hb.size(), cudaMemcpyDeviceToHost);
    cudaFree(ud_dev);
    if (FILE* f = fopen(getenv("DPI_TC_UD_FILE"), "wb")) {
      fwrite(hb.data(), sizeof(float2), hb.size(), f);
      fclose(f);
    }
  }
#endif
  if (N.nchunks > 1)
    cha_tc_finish_multi_kernel<<<(unsigned)n, 128, 0, st>>>(partG, partF, A.units_per_point, P.nx, M_I, bound, y, nonfinite);
  else
    cha_tc_finish_kernel<<<(unsigned)n, 128, 0, st>>>(partG, partF, A.units_per_point, N.cls, P.nx, M_I, bound, y, nonfinite);
  DPI_CUDA_CHECK(cudaGetLastError());
  return DPI_OK;
}

int tc_umma_selftest(const float* A, const float* B, float* D, cudaStream_t st) {
  uint8_t* img = nullptr;
  DPI_CUDA_CHECK(cudaMalloc((void**)&img, 2 * (size_t)tc::kOperandBytes));
  pack_sw128_kernel<<<64, 256, 0, st>>>(A, 128, 0, 128, 128, img);
  pack_sw128_kernel<<<64, 256, 0, st>>>(B, 128, 0, 128, 128, img + tc::kOperandBytes);
  const int smem = 2 * tc::kOperandBytes + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) {
    const char* m = getenv("DPI_SELFTEST_MODE");  // developer switch: see the kernel comment
    umma_selftest_kernel<<<1, 128, smem, st>>>(img, img + tc::kOperandBytes, D, m ? atoi(m) : 0);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(img);
  if (e != cudaSuccess) {
    set_error("umma selftest failed: %s", cudaGetErrorString(e));
    return DPI_ERR_CUDA;
  }
  return DPI_OK;
}

}  // namespace dpi
