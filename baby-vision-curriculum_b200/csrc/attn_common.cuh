// attn_common.cuh -- tile constants, TMA / TMEM helpers and operand descriptors shared by the attention kernels
// (attn.cu: any sequence length; attn_small.cu: whole-sequence-resident kernels for S <= 192).
#pragma once
#include "../../include/bvc.h"
#include "bvc_host.h"
#include "bvc_ptx.cuh"

namespace bvc {

constexpr int kTile = 128;          // rows per Q / KV tile
constexpr int kTileBytes = 16384;   // 128 x 64 bf16
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// Placed right after tcgen05.wait::ld: names the loaded registers as read-write operands of (empty) volatile asm
// statements, which keep their order after the wait -- so nothing that consumes them can be scheduled above it.
template <int N>
__device__ __forceinline__ void tmem_ld_pin(uint32_t (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major operand tile [rows][64]: k-step (16 elements) = +32 bytes inside the swizzle atom
__device__ __forceinline__ uint64_t desc_k(uint32_t base, int kstep) { return umma_smem_desc(base + kstep * 32, 1024, 16); }
// MN-major operand tile [k rows][64 mn]: k-step (16 rows) = +2048 bytes; lbo = distance between 64-wide mn blocks
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int kstep, uint32_t lbo) {
  return umma_smem_desc(base + kstep * 2048, 1024, lbo);
}
// byte offset of 8 consecutive bf16 (one 16-byte chunk) of element (row, col8*8) in a [128][128] bf16 tile stored as
// two [128 rows][64] 128B-swizzled sub-tiles (the layout TMA would have produced)
__device__ __forceinline__ uint32_t ptile_chunk_off(int row, int chunk16 /*0..15*/) {
  return (uint32_t)((chunk16 >> 3) * kTileBytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4));
}


// [B, S, heads_total, 64] bf16 viewed as a 4-D tensor; box = 64 x 1 x 128 x 1 -> one [128 rows][64] tile
static inline int make_head_tmap(CUtensorMap* tm, const void* base, int heads_total, int S, int B, int box_rows = 128) {
  const uint64_t dims[4] = {64, (uint64_t)heads_total, (uint64_t)S, (uint64_t)B};
  const uint64_t strides[3] = {128, (uint64_t)heads_total * 128, (uint64_t)S * heads_total * 128};
  const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace bvc
