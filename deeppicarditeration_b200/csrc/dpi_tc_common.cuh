// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA bulk copy, tcgen05 (alloc / mma / commit / ld),
// UMMA shared-memory descriptors and the SWIZZLE_128B K-major operand layout.  Inline PTX only (no CUTLASS).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpi {
namespace tc {

// One [128 rows x 128 K] bf16 operand in the canonical K-major SWIZZLE_128B layout = two 64-column K panels of
// 128 rows x 128 bytes; inside a panel the 16-byte chunk index of a row is XORed with (row & 7).
constexpr int kPanelBytes = 128 * 128;          // 16 KB
constexpr int kOperandBytes = 2 * kPanelBytes;  // 32 KB

// byte offset of element (row, k) inside an operand image (row < 128, k < 128)
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
  const uint32_t panel = k >> 6, chunk = (k & 63u) >> 3;
  return panel * kPanelBytes + row * 128u + (((chunk ^ (row & 7u)) << 4) | ((k & 7u) << 1));
}

// instruction descriptor, kind::f16: D=f32 (bit 4), A=bf16 (bit 7), B=bf16 (bit 10), both K-major,
// N>>3 at bits [17,23), M>>4 at bits [24,29)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same with fp16 operands (format code 0 for A and B)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same with fp16 ACCUMULATORS (c_format = 0): one half per 32-bit TMEM cell, read back two columns per register by
// tcgen05.ld ... .pack::16b (checked by the UMMA self-test, modes 1 and 2: 4.4e-4 from the fp32-accumulated product
// at K = 128, against 2.8e-4 for rounding an fp32 accumulator to fp16 -- which every epilogue pass does anyway)
__host__ __device__ constexpr uint32_t make_idesc_f16_acc16(int M, int N) {
  return ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::tf32: fp32 words in shared memory read as TF32 (format code 2 for A and B), K = 8 per instruction
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor for a K-major SWIZZLE_128B operand whose 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset: unused for swizzled K-major (set to 1)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset = 8 rows * 128 B   bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}
// descriptor of the UMMA_K=16 slice `ks` (0..7) of an operand image at `base`
__device__ __forceinline__ uint64_t desc_kstep(uint32_t base, int ks) {
  return make_desc_sw128(base + (uint32_t)(ks >> 2) * kPanelBytes + (uint32_t)(ks & 3) * 32u);
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the warp sleeps in hardware until the phase completes or the hint
// expires, instead of burning issue slots that the epilogue and producer warps of the same SMSP need.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a protocol bug must trap (and fail the launch) instead of hanging the GPU box.  The slow
// path reads the nanosecond timer once on entry and then only every 64th retry (the first version read clock64 on
// every retry and the waiting warps issued a third of all instructions of the kernel).
#ifndef DPI_TC_WAIT_HINT_NS
#define DPI_TC_WAIT_HINT_NS 20000u
#endif
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifndef DPI_TC_SLIM_WAIT
#define DPI_TC_SLIM_WAIT 1
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
#if DPI_TC_SLIM_WAIT
  // One asm loop: try_wait (hardware sleep up to the hint, woken by barrier traffic), retry counter, trap.  The
  // retry path is 5 instructions; the C++ loop with a timer read was 17, and ncu's source view attributed 45 % of
  // all executed warp instructions of the integral kernel to the waiting warps -- issue slots taken from the
  // epilogue warps of the same sub-partition.  2^22 retries of 60 ns (woken early) to 20 us (the hint) each: the watchdog fires after 0.25 .. 84 s.
  (void)tag;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 c;\n"
      "mov.u32 c, 0;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra.uni DONE_%=;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 q, c, 0x400000;\n"
      "@q bra.uni WAIT_%=;\n"
      "trap;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(DPI_TC_WAIT_HINT_NS)
      : "memory");
#else
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, DPI_TC_WAIT_HINT_NS)) {
    if ((++spins & 63u) == 0 && global_ns() - t0 > 2000000000ull) {  // 2 s: three orders above any legitimate wait
#ifdef DPI_TC_WATCHDOG_PRINTF
      printf("dpi tc watchdog: block %d thread %d stuck on barrier tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
#endif
      __trap();
    }
  }
#endif
}
// Wait of a warp that runs well AHEAD of its consumer (ring producers): poll at most every ~0.4 us.  mbar_wait's
// try_wait is woken by every barrier event of the CTA (measured: one retry per ~150 cycles per waiting warp, 7
// instructions each -- a fifth of all issued instructions of the SDGD kernel); a producer with rows of slack does
// not need that reaction time.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(400);
    if (++spins > (1u << 23)) __trap();  // > 3 s
  }
}

// 16-byte shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, one UMMA_K=16 step, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent issue: every lane executes the surrounding (uniform) address arithmetic, so the descriptors live in
// uniform registers, and only the elected lane issues.  (Issuing from inside an `if (lane == 0)` region makes the
// compiler wrap every tcgen05.mma in a waterfall loop of R2UR broadcasts, ~150 cycles per instruction.)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t p;
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(p));
  return p;
}
__device__ __forceinline__ void umma_bf16_p(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_p(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tc_commit_p(uint32_t bar, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar),
      "r"(leader)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_p(uint32_t bar, uint32_t bytes, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n}" ::"r"(bar),
      "r"(bytes), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_p(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar,
                                           uint32_t leader) {
  asm volatile(
      "{\n.reg .pred q;\nsetp.ne.b32 q, %4, 0;\n"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "r"(leader)
      : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of warp w reads TMEM lane 32*(w%4)+i, columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 32 columns of 16-bit accumulators (c_format = F16: one half per 32-bit TMEM cell), adjacent columns packed
// two per register: thread i of warp w reads lane 32*(w%4)+i, columns [c, c+32) -> r[j] = {col c+2j, col c+2j+1}
__device__ __forceinline__ void tmem_ld16_pack(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_pack(uint32_t taddr, uint32_t (&r)[8]) {  // 16 columns -> 8 registers
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld4_pack(uint32_t taddr, uint32_t (&r)[4]) {  // 8 columns -> 4 registers
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.pack::16b.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// ---- packed fp16x2 arithmetic (two columns per instruction) ---------------------------------------
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t h2_abs(uint32_t a) { return a & 0x7FFF7FFFu; }
// max(-a, 0) per half
__device__ __forceinline__ uint32_t h2_negrelu(uint32_t a) {
  uint32_t r;
  asm("{\n.reg .b32 n;\nneg.f16x2 n, %1;\nmax.f16x2 %0, n, %2;\n}" : "=r"(r) : "r"(a), "r"(0u));
  return r;
}
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_ex2(uint32_t a) {
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
// 2^x for x <= 0, packed fp16x2, WITHOUT the MUFU pipe: n = round(-x) by the magic-number trick (1536 + (-x) has
// ulp 1 in fp16), f = x + n in [-1/2, 1/2], 2^f by a cubic in Horner form (max rel error 1.0e-4 in exact arithmetic;
// 6.2e-4 max / 1.9e-4 mean evaluated in fp16, against 4.8e-4 / 1.7e-4 for a correctly rounded fp16 result), and the
// scale 2^-n applied by subtracting n from the exponent fields of both halves with one 32-bit integer subtract
// (2^f >= 0.707 has exponent field >= 14 and n <= 14, so no borrow crosses the halves).  x is clamped to -14.
// 10 FMA/ALU-pipe instructions per pair instead of 2 MUFU + 1 PRMT: the epilogue passes are bound by the quarter-rate
// XU pipe (ex2 and the f32 -> f16x2 packs), not by issue slots, so part of the columns take this route.
__device__ __forceinline__ uint32_t h2_ex2_poly(uint32_t x) {
  constexpr uint32_t kMagic = 0x66006600u, kNegMagic = 0xE600E600u, kNeg1 = 0xBC00BC00u, kClamp = 0xCB00CB00u;  // 1536, -1536, -1, -14
  constexpr uint32_t c0 = 0x3C003C00u, c1 = 0x398C398Cu, c2 = 0x33C433C4u, c3 = 0x2B282B28u;  // 1, 0.6934, 0.2427, 0.0559
  x = h2_max(x, kClamp);
  const uint32_t t = h2_fma(x, kNeg1, kMagic);
  const uint32_t f = h2_add(x, h2_add(t, kNegMagic));
  uint32_t p = h2_fma(f, c3, c2);
  p = h2_fma(p, f, c1);
  p = h2_fma(p, f, c0);
  return p - ((t << 10) & 0xFC00FC00u);
}
__device__ __forceinline__ float2 h2_unpack(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__host__ __device__ __forceinline__ uint32_t h2_const(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace tc
}  // namespace dpi
