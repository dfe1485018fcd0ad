// Shared device helpers for the gdmcf sm_100a kernels: PTX wrappers for mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and a Philox-4x32-10 counter RNG.
// Everything here is sm_100a-only; there is deliberately no fallback path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define GD_DEV __device__ __forceinline__

namespace gd {

GD_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

GD_DEV uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Programmatic dependent launch (see api_internal.h): let the next kernel's CTAs be scheduled now, then wait until
// every kernel this one depends on has completed and its writes are visible. Both are no-ops for a kernel that was
// launched without the attribute / has no programmatic dependents.
GD_DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
GD_DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
GD_DEV void pdl_entry() {
  pdl_launch_dependents();
  pdl_wait();
}

GD_DEV bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
GD_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
GD_DEV void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GD_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
GD_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
GD_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> CUDA error -> rc -3), never hang the GPU.
GD_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == 4096u) t0 = globaltimer_ns();
    if (spins > 4096u && (spins & 1023u) == 0u) {
      if (globaltimer_ns() - t0 > 4000000000ull) __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA: 2-D tiled bulk tensor load, completion on an mbarrier (complete_tx::bytes)
// ----------------------------------------------------------------------------------------------
GD_DEV void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
GD_DEV void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}

// 2-D tiled bulk tensor store smem -> global (bulk async-group completion). Out-of-bounds parts of the box are
// clipped by the hardware, so M/N tails need no masking.
GD_DEV void tma_store_2d(const void* tmap, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tmap),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
GD_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest 0 groups have finished READING their shared-memory source (staging can be overwritten)
GD_DEV void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
GD_DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA) that will read them
GD_DEV void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
GD_DEV void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ----------------------------------------------------------------------------------------------
// tcgen05: tensor memory + 5th-gen tensor-core MMA (single-CTA group)
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
GD_DEV void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
GD_DEV void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
GD_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
GD_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulation.
GD_DEV void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired.
GD_DEV void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster (same TPC) run one 256-row MMA; each holds 128 rows of A, half of the
// B tile and its 128 accumulator rows. One thread of the leader (cluster rank 0) issues the MMAs for both.
// ----------------------------------------------------------------------------------------------
GD_DEV uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
GD_DEV void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
GD_DEV uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
GD_DEV void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
GD_DEV void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
GD_DEV void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are counted on an mbarrier that may live in the
// peer CTA of the pair (`bar_cluster_addr` is a shared::cluster address, see mapa_shared).
GD_DEV void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
      : "memory");
}
GD_DEV void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this offset in every CTA of `cta_mask` once all MMAs issued so far have retired.
GD_DEV void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile laid out by TMA with SWIZZLE_128B:
// rows are 128 B (64 bf16) apart, 8-row swizzle atoms are 1024 B apart (SBO), LBO unused (=1).
// Field layout follows the sm_100 "matrix descriptor" (start>>4 @0, LBO>>4 @16, SBO>>4 @32,
// version=1 @46, layout SWIZZLE_128B=2 @61).
GD_DEV uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both operands K-major (bits 15,16 = 0), N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
GD_DEV void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
GD_DEV void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// Philox-4x32-10 (counter-based; same construction as cuRAND's / torch's generator family).
// ----------------------------------------------------------------------------------------------
struct Philox {
  uint32_t key0, key1;
  GD_DEV Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
  GD_DEV uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
    uint32_t k0 = key0, k1 = key1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
GD_DEV float u32_to_unit_open(uint32_t x) {  // (0,1]
  return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}
GD_DEV float2 box_muller(uint32_t a, uint32_t b) {
  float u = u32_to_unit_open(a);
  float v = u32_to_unit_open(b);
  float r = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * v, &s, &c);
  return make_float2(r * c, r * s);
}

// One AdamW element update (torch.optim.AdamW, amsgrad=False, maximize=False) with every rounding spelled out, so that
// the flat kernel and the fused update+refresh kernels produce bit-identical parameters whatever the compiler would
// otherwise contract into FMAs.
struct AdamwCoef {
  float decay, one_m_b1, beta2, one_m_b2, eps, step_size, bc2_sqrt, gscale;
};
GD_DEV void adamw_update(const AdamwCoef& k, float& param, float gr, float& mi, float& vi) {
  const float grad = __fmul_rn(gr, k.gscale);
  param = __fmul_rn(param, k.decay);
  mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(grad, mi), k.one_m_b1));                          // exp_avg.lerp_(grad, 1 - beta1)
  vi = __fadd_rn(__fmul_rn(vi, k.beta2), __fmul_rn(__fmul_rn(k.one_m_b2, grad), grad));    // mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), k.bc2_sqrt), k.eps);
  param = __fsub_rn(param, __fmul_rn(k.step_size, __fdiv_rn(mi, denom)));
}
GD_DEV AdamwCoef adamw_coef(float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                            float grad_scale, const long long* step_dev) {
  if (step_dev) {  // bias corrections from a device-resident step counter (CUDA-graph replays advance it on the device)
    const double st = (double)step_dev[0];
    bc1 = (float)(1.0 - pow((double)beta1, st));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, st));
  }
  AdamwCoef k;
  k.decay = 1.0f - lr * weight_decay;
  k.one_m_b1 = 1.0f - beta1;
  k.beta2 = beta2;
  k.one_m_b2 = 1.0f - beta2;
  k.eps = eps;
  k.step_size = lr / bc1;
  k.bc2_sqrt = bc2_sqrt;
  k.gscale = grad_scale;
  return k;
}

GD_DEV float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace gd
