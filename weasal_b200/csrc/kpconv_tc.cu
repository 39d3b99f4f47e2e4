// KPConv forward / backward for sm_100a: sparse kernel-point gather on CUDA cores feeding tcgen05 (TF32 in, FP32
// accumulate in TMEM) through a warp-specialised, multi-stage shared-memory pipeline — replaces models/blocks.py:238-374
// (rigid, 'linear' influence, 'sum' aggregation) and the autograd backward of that expression.
//
//   out[i,:] = sum_k ( sum_h w[i,k,h] * x[idx[i,h],:] ) @ W[k],   w = max(0, 1 - ||(s[idx[i,h]] - q[i]) - kp[k]|| / ext)
//
// Facts the design rests on (measured on ALS spheres):
//   * w is ~93 % zeros: a neighbour lies within KP_extent of ~1.03 kernel points (max 3). So the first contraction
//     is done sparsely in fp32 on CUDA cores (exact, ~15x fewer FMAs than the dense [K x H] x [H x Cin] product),
//     and only the dense second contraction [P x (K*Cin)] x [(K*Cin) x Cout] goes to the tensor cores.
//   * bf16 operands give ~1.6e-3 relative error on that contraction, above the 1e-3 parity bar; TF32 operands
//     rounded to nearest give ~4e-4. Hence kind::tf32.
//   * the sparse gather reads Nq*H_real*Cin*4 bytes out of L2 (~4-5x the bytes the tensor core consumes), so the kernel
//     is bound by how many gathers the producer warps keep in flight; the pipeline exists to keep them issuing.
//
// Kernels
//   kp_lists       one CTA per tile of 128 centre points, ALL list jobs of a call in one grid: a per-CTA candidate table
//                  (8^3 cells -> kernel points that can reach the cell) cuts the influence evaluations to ~3 per
//                  neighbour; hits are compacted into ONE FLAT LIST PER (tile, kernel point), sorted by row: entry =
//                  (neighbour index | row in tile << 25, weight). A tile header holds, per kernel point, the list's start
//                  at every 8-row block. The lists depend on the geometry and the (frozen) kernel points only: a
//                  training step builds them in its prefetch stage, off the training stream.
//   kp_pack_w      W[k,c,o] -> TF32-rounded B-operand images, one per 64-column chunk of the (k,c) reduction axis,
//                  already in the UMMA K-major core-matrix layout (a CTA fetches a chunk with one bulk copy); all
//                  layers of a step in one launch.
//   kp_fwd         one CTA per 128-point tile (x a slice of the reduction axis on deep layers). Warp roles: 8 PRODUCER
//                  warps per stage assemble A chunks [128 x 64] into a ring of shared-memory stages (each lane group
//                  walks one flat list segment: U independent float4 gathers in flight, accumulation in registers, one
//                  store per row; with one CTA per SM two groups of 8 warps fill alternate stages); 1 LOADER warp streams
//                  the packed weight chunks with cp.async.bulk into its own ring (opt-in: multicast inside a cluster);
//                  1 MMA warp issues tcgen05.mma.kind::tf32 (M128 x N x K8) into TMEM and releases stages with
//                  tcgen05.commit; the producer warps drain TMEM at the end (fused bias / LeakyReLU / split-reduction
//                  atomics). Backward-dX is the same kernel run on the transposed neighbour table with W^T and -kp
//                  (atomics-free segmented scatter).
//   kp_dw          dW[(k,c),o] = sum_i WF[i,(k,c)] * dOut[i,o]: stages of 64 points, the A tile consumed MN-major
//                  (M = (k,c) rows, K = points) against the dOut tile, which the loader warp fetches with TMA tensor
//                  copies (cp.async.bulk.tensor.2d, TFLOAT32 conversion and 32-byte-atom swizzle done by the copy
//                  engine); same producer / MMA roles, accumulated in TMEM across a CTA's point tiles, then added to dW.
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace kp {

// ------------------------------------------------------------------------------------------------------- constants
constexpr int TILE_M = 128;       // centre points per tile (= UMMA M of the forward / dX kernels)
constexpr int RB = 8;             // rows per row block (= one core-matrix row group)
constexpr int NRB = TILE_M / RB;  // 16 row blocks per tile
constexpr int TOFF_RB = NRB + 1;  // per kernel point: list position at the start of every row block, plus its end
constexpr int TOFF_PER_TILE = 16 * TOFF_RB;   // 272 ints per tile; slot [15][0] = the tile's first entry, [15][1] = count
constexpr int ROW_SHIFT = 25;     // entry.x = neighbour index | (row in tile << 25)
constexpr unsigned J_MASK = (1u << ROW_SHIFT) - 1u;
constexpr int K_SHIFT = 27;       // scratch lists (per row, grouped by kernel point): index | (kernel point << 27)
constexpr unsigned JS_MASK = (1u << K_SHIFT) - 1u;

// Producer warps per CTA (template parameter NPW of the kernels; they are also the epilogue warps): 8 where two CTAs
// share an SM, 16 where the stages of a wide layer leave room for one CTA only. A stage is always filled by 8 warps
// (every lane owns 8 rows of one column group); with 16 warps the two groups of 8 fill alternate stages. The gather is
// latency bound, so an SM wants ~16 producer warps either way. Two more warps follow them: the loader and the MMA issuer.
constexpr int FWD_CK = 64;        // reduction columns per forward stage
constexpr int DW_CK = 128;        // reduction columns (= UMMA M) per dW CTA
constexpr int DW_PT = 64;         // points per dW stage
constexpr int MAX_STAGES = 4;
#ifndef KP_U
#define KP_U 4            // entries (gathers) in flight per producer lane
#endif

// forward A stage [128 rows x 64 cols], UMMA canonical K-major no-swizzle layout: element (row p, col c) at
//   (c/4)*A_LBO + (p/8)*A_SBO + (p%8)*16 + (c%4)*4     (8 rows x 16 bytes core matrices)
// A_LBO carries 16 bytes of padding so that lanes writing one row are spread over the banks.
constexpr int A_SBO = 128;
constexpr int A_LBO = 16 * 128 + 16;              // 2064
constexpr int A_STAGE = (FWD_CK / 4) * A_LBO;     // 33024
constexpr int B_SBO = 128;
// dW stages, MN-major SWIZZLE_128B_BASE32B (the only MN-major layout tcgen05 accepts for 32-bit operands,
// cute::UMMA::Layout_MN_SW128_32B_Atom): rows of 32 elements (128 B) along M/N, 4 consecutive K values (points) = 4
// consecutive rows (512 B atom), byte-address bits [5,7) XORed with bits [7,9). Element (m, k) at
//   (m/32)*MN_LBO + (k/4)*MN_SBO + (k%4)*128 + ((((m%32)/8) ^ (k%4))*32) + (m%8)*4
constexpr int MN_SBO = 512;
constexpr int MN_LBO = (DW_PT / 4) * MN_SBO;      // 8 KiB per group of 32 M/N values
constexpr int DW_A_STAGE = (DW_CK / 32) * MN_LBO; // 32 KiB

// --------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out: neither a spin
// (a plain try_wait loop returned every ~20 ns and took 37 % of the kernel's issue slots away from the producers, ncu
// capture profiles/r2_ncu_ws_v2_spin.csv) nor a nanosleep back-off (which adds its quantum to every hand-over of the ring).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(4000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded: a barrier that never completes (a malformed descriptor, a lost bulk copy) traps instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// thread-block clusters: a weight chunk is fetched ONCE per cluster, every CTA copies its slice into the shared memory
// of all CTAs of the cluster (multicast) and signals each CTA's own barrier at the same offset
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {  // arrives on `bar` of every CTA in the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// 2-D tiled tensor copy (TMA): box of the tensor map at (c0 = innermost coordinate, c1) -> shared memory
__device__ __forceinline__ void tma_g2s_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` when every MMA issued so far is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {  // 32 lanes x 16 columns, one row per thread
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float to_tf32(float f) {  // round to nearest, ties away (the MMA itself truncates)
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 to_tf32(float4 v) {
    return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// [0,14) start>>4, [16,30) leading byte offset>>4, [32,46) stride byte offset>>4, [61,64) layout type
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 0) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;  // cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ influence lists
struct Table {           // neighbour table of the centre points
    const void* idx;     // padded rows [nc, H] (row stride `stride`), or CSR column indices when rowptr != null
    const int* rowptr;   // CSR row pointers [nc+1] or null
    int H, stride, is_i64;
};

__device__ __forceinline__ long long table_get(const Table& T, size_t pos) {
    return T.is_i64 ? ((const long long*)T.idx)[pos] : (long long)((const int*)T.idx)[pos];
}

__device__ __forceinline__ float influence_w(float rx, float ry, float rz, float kx, float ky, float kz, float inv_ext) {
    const float dx = rx - kx, dy = ry - ky, dz = rz - kz;
    return fmaxf(0.f, 1.f - sqrtf(dx * dx + dy * dy + dz * dz) * inv_ext);
}

constexpr int LST_WARPS = 16;     // warps per list-building CTA: 8 rows of the tile each
constexpr int LST_ROWS = TILE_M / LST_WARPS;
constexpr int LST_G = 4;          // rows a warp has in flight
constexpr int LST_CELLS = 8;      // candidate table: LST_CELLS^3 cells over the cube that holds every kernel point's ball

// One CTA per tile of 128 centres.
//   Candidate table (per CTA, 512 cells, a few thousand instructions): which kernel points can reach a relative position,
//     by cell of a coarse grid over the cube [-b, b]^3, b = max |kernel point coordinate| + extent (conservative
//     point-to-box test). A neighbour lies within the extent of ~1.03 kernel points and its cell names ~3 candidates, so
//     the exact influence (the same expression as before: the set of entries and their weights do not change) is
//     evaluated ~3 times per neighbour instead of 15.
//   Phase A (one warp per centre at a time, 8 centres per warp; lanes = columns of the neighbour table): every lane looks
//     its neighbour's candidates up and the warp walks them, one candidate per lane per pass; non-zero weights are ballot-
//     compacted into the row's slot of a scratch buffer (15 entries per table cell is the exact worst case, so no
//     allocation), tagged with the kernel point, and counted per (row, kernel point) in shared memory.
//   Scan: per kernel point, the exclusive prefix of the counts over the 128 rows = each row's place in the tile's flat
//     list of that kernel point; the tile takes its range of the compact entry buffer with one atomicAdd.
//   Phase B: the rows' entries are copied from the scratch slots (still in L2) to their place in the flat lists; the rank
//     of an entry among the row's entries of the same kernel point comes from __match_any_sync (deterministic order).
struct ListsOut {
    int* toff;        // [n_tiles][16][17]
    int2* entries;    // compact, `cap` entries
    int* ctl;         // [0] = entries used so far (atomic cursor), [1] = overflow flag
    long long cap;
    int* overflow;    // optional: one flag shared by all lists of a batch (set together with ctl[1])
};

// One launch builds the lists of several (centres, table, kernel points) jobs: the 20 lists of a training step's 10
// KPConv (forward + dX each) are one grid of ~1700 tiles instead of 20 launches that each end in a partial wave (the
// deep layers have 4-16 tiles).
struct ListJob {
    const float* centres;
    const float* others;
    Table T;
    const float* kp;
    int2* scratch;
    ListsOut L;
    int nc, no, K;
    float kp_sign, inv_ext;
    int first_tile;   // of this job inside the grid
};
constexpr int LST_MAX_JOBS = 24;
struct ListJobs {
    ListJob job[LST_MAX_JOBS];
    int n;
};

__global__ void __launch_bounds__(128) kp_lists_ctl_kernel(const __grid_constant__ ListJobs J) {
    if ((int)threadIdx.x < 4 * J.n)  // the 4 control ints of every job's header (cursor, overflow flag, 2 spare)
        J.job[threadIdx.x >> 2].L.ctl[threadIdx.x & 3] = 0;
}

__global__ void __launch_bounds__(LST_WARPS * 32, 2) kp_lists_kernel(const __grid_constant__ ListJobs J) {
    int ji = 0;
    while (ji + 1 < J.n && (int)blockIdx.x >= J.job[ji + 1].first_tile) ji++;
    const float* __restrict__ centres = J.job[ji].centres;
    const float* __restrict__ others = J.job[ji].others;
    const Table T = J.job[ji].T;
    const float* __restrict__ kp = J.job[ji].kp;
    int2* __restrict__ scratch = J.job[ji].scratch;
    const ListsOut L = J.job[ji].L;
    const int nc = J.job[ji].nc, no = J.job[ji].no, K = J.job[ji].K;
    const float kp_sign = J.job[ji].kp_sign, inv_ext = J.job[ji].inv_ext;
    const int tile = (int)blockIdx.x - J.job[ji].first_tile;
    constexpr int NCELL = LST_CELLS * LST_CELLS * LST_CELLS;
    __shared__ float4 s_kp[16];
    __shared__ unsigned short s_tab[NCELL];        // bit k: kernel point k may influence positions of the cell
    __shared__ unsigned short s_cnt[TILE_M][16];   // per row: entries per kernel point; [15] = the row's total
    __shared__ int s_pos[15][TILE_M + 1];          // per kernel point: exclusive prefix of the counts over the rows
    __shared__ int s_start[17];                    // list starts inside the tile's range; [15] = tile total; [16] = range base
    if (threadIdx.x < 16) {
        const int k = threadIdx.x;
        s_kp[k] = k < K ? make_float4(kp_sign * kp[3 * k], kp_sign * kp[3 * k + 1], kp_sign * kp[3 * k + 2], 0.f)
                        : make_float4(1e30f, 1e30f, 1e30f, 0.f);
    }
    for (int t = threadIdx.x; t < TILE_M * 16 / 2; t += LST_WARPS * 32) reinterpret_cast<unsigned*>(&s_cnt[0][0])[t] = 0u;
    __syncthreads();
    const float ext = 1.f / inv_ext;
    float box = 0.f;
    for (int k = 0; k < K; k++) box = fmaxf(box, fmaxf(fabsf(s_kp[k].x), fmaxf(fabsf(s_kp[k].y), fabsf(s_kp[k].z))));
    box += ext;
    const float cell = 2.f * box / LST_CELLS, inv_cell = LST_CELLS / (2.f * box);
    {
        // the half edge is padded: a position within rounding distance of a cell face may be binned on either side
        const float half = 0.5f * cell * 1.001f + 1e-6f * box, reach2 = ext * ext * 1.0001f;
        for (int c = threadIdx.x; c < NCELL; c += LST_WARPS * 32) {
            const int ix = c % LST_CELLS, iy = (c / LST_CELLS) % LST_CELLS, iz = c / (LST_CELLS * LST_CELLS);
            const float ccx = -box + (ix + 0.5f) * cell, ccy = -box + (iy + 0.5f) * cell, ccz = -box + (iz + 0.5f) * cell;
            unsigned m = 0;
            for (int k = 0; k < K; k++) {
                const float4 q = s_kp[k];
                const float dx = fmaxf(0.f, fabsf(q.x - ccx) - half), dy = fmaxf(0.f, fabsf(q.y - ccy) - half),
                            dz = fmaxf(0.f, fabsf(q.z - ccz) - half);
                if (dx * dx + dy * dy + dz * dz < reach2) m |= 1u << k;
            }
            s_tab[c] = (unsigned short)m;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;

    // rows are taken LST_G at a time: the table reads and the coordinate gathers of the group are issued together (the
    // kernel is bound by the latency of the chain table entry -> coordinates, not by its instruction count)
    for (int g0 = 0; g0 < LST_ROWS; g0 += LST_G) {
        const int row_g = warp * LST_ROWS + g0;
        int cnt_r[LST_G], run[LST_G];
        size_t pos0[LST_G], row0[LST_G];
        float cx[LST_G], cy[LST_G], cz[LST_G];
        int max_cnt = 0;
#pragma unroll
        for (int rr = 0; rr < LST_G; rr++) {
            const int i = tile * TILE_M + row_g + rr;
            cnt_r[rr] = 0; run[rr] = 0; pos0[rr] = 0; row0[rr] = 0;
            cx[rr] = cy[rr] = cz[rr] = 0.f;
            if (i < nc) {
                cx[rr] = centres[3 * (size_t)i]; cy[rr] = centres[3 * (size_t)i + 1]; cz[rr] = centres[3 * (size_t)i + 2];
                if (T.rowptr) { row0[rr] = (size_t)T.rowptr[i]; pos0[rr] = row0[rr]; cnt_r[rr] = T.rowptr[i + 1] - T.rowptr[i]; }
                else { row0[rr] = (size_t)i * T.H; pos0[rr] = (size_t)i * T.stride; cnt_r[rr] = T.H; }
            }
            max_cnt = max(max_cnt, cnt_r[rr]);
        }
        for (int hb = 0; hb < max_cnt; hb += 32) {
            const int h = hb + lane;
            int jv[LST_G];
            float rx[LST_G], ry[LST_G], rz[LST_G];
#pragma unroll
            for (int rr = 0; rr < LST_G; rr++) {
                const long long j = (h < cnt_r[rr]) ? table_get(T, pos0[rr] + h) : -1;
                jv[rr] = (j >= 0 && j < no) ? (int)j : -1;
            }
#pragma unroll
            for (int rr = 0; rr < LST_G; rr++) {
                rx[rr] = ry[rr] = rz[rr] = 0.f;
                if (jv[rr] >= 0) {
                    rx[rr] = others[3 * (size_t)jv[rr]] - cx[rr]; ry[rr] = others[3 * (size_t)jv[rr] + 1] - cy[rr];
                    rz[rr] = others[3 * (size_t)jv[rr] + 2] - cz[rr];
                }
            }
#pragma unroll
            for (int rr = 0; rr < LST_G; rr++) {
                unsigned short* cnt = s_cnt[row_g + rr];
                int2* my_entries = scratch + 15 * row0[rr];
                unsigned cand = 0;
                if (jv[rr] >= 0) {
                    const float fx = (rx[rr] + box) * inv_cell, fy = (ry[rr] + box) * inv_cell, fz = (rz[rr] + box) * inv_cell;
                    if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)LST_CELLS && fy < (float)LST_CELLS && fz < (float)LST_CELLS)
                        cand = s_tab[((int)fz * LST_CELLS + (int)fy) * LST_CELLS + (int)fx];
                }
                while (__any_sync(0xffffffffu, cand != 0u)) {
                    const bool act = cand != 0u;
                    const int k = act ? __ffs((int)cand) - 1 : 0;
                    cand &= cand - 1u;
                    float w = 0.f;
                    if (act) {
                        const float4 q = s_kp[k];
                        w = influence_w(rx[rr], ry[rr], rz[rr], q.x, q.y, q.z, inv_ext);
                    }
                    const bool hit = w > 0.f;
                    const unsigned hm = __ballot_sync(0xffffffffu, hit);
                    if (hm == 0u) continue;
                    if (hit) {
                        const unsigned peers = __match_any_sync(hm, k);
                        int2 e;
                        e.x = (int)((unsigned)jv[rr] | ((unsigned)k << K_SHIFT));
                        e.y = __float_as_int(w);
                        my_entries[run[rr] + __popc(hm & lt_mask)] = e;
                        if ((peers & lt_mask) == 0u) cnt[k] = (unsigned short)(cnt[k] + __popc(peers));
                    }
                    run[rr] += __popc(hm);
                    __syncwarp();  // the next pass may count the same kernel point from another lane
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < LST_G; rr++)
            if (lane == 0 && tile * TILE_M + row_g + rr < nc) s_cnt[row_g + rr][15] = (unsigned short)(run[rr] < 65535 ? run[rr] : 65535);
    }
    __syncthreads();

    // per kernel point: exclusive prefix over the rows (warp w takes kernel points w, w + 16, ...)
    for (int k = warp; k < 15; k += LST_WARPS) {
        int carry = 0;
        for (int r0 = 0; r0 < TILE_M; r0 += 32) {
            const int r = r0 + lane;
            const int c = (int)s_cnt[r][k];
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_pos[k][r] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_pos[k][TILE_M] = carry;
    }
    __syncthreads();
    if (warp == 0) {
        const int len = lane < 15 ? s_pos[lane][TILE_M] : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < 16) s_start[lane] = incl - len;  // [15] = total
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 0) {
            base = total > 0 ? atomicAdd(L.ctl, total) : 0;
            if ((long long)base + total > L.cap) {  // (only possible for a caller-bounded buffer: the batch takes the
                atomicExch(L.ctl + 1, 1);           //  slow path, the tile is left empty)
                if (L.overflow) atomicExch(L.overflow, 1);
                base = -1;
            }
            s_start[16] = base;
        }
    }
    __syncthreads();
    const int base = s_start[16];
    int* toff = L.toff + (size_t)tile * TOFF_PER_TILE;
    for (int t = threadIdx.x; t < TOFF_PER_TILE; t += LST_WARPS * 32) {
        const int k = t / TOFF_RB, rb = t - k * TOFF_RB;
        int v;
        if (k < 15) v = base < 0 ? 0 : s_start[k] + s_pos[k][rb * RB];
        else v = rb == 0 ? (base < 0 ? 0 : base) : (rb == 1 ? (base < 0 ? 0 : s_start[15]) : 0);
        toff[t] = v;
    }
    if (base < 0) return;
    int2* dst = L.entries + base;
    // the first 32 entries of every row of a group are read together (most rows have fewer), the rest row by row
    for (int g0 = 0; g0 < LST_ROWS; g0 += LST_G) {
        const int row_g = warp * LST_ROWS + g0;
        int total[LST_G];
        const int2* src[LST_G];
        int2 first[LST_G];
#pragma unroll
        for (int rr = 0; rr < LST_G; rr++) {
            const int row = row_g + rr, i = tile * TILE_M + row;
            total[rr] = 0;
            src[rr] = scratch;
            if (i < nc) {
                total[rr] = s_cnt[row][15];
                if (total[rr] == 65535) {  // (a row of more than 4369 table cells: recount)
                    total[rr] = 0;
                    for (int k = 0; k < 15; k++) total[rr] += (int)s_cnt[row][k];
                }
                src[rr] = scratch + 15 * (T.rowptr ? (size_t)T.rowptr[i] : (size_t)i * T.H);
            }
            first[rr] = make_int2(0, 0);
            if (lane < total[rr]) first[rr] = src[rr][lane];
        }
#pragma unroll
        for (int rr = 0; rr < LST_G; rr++) {
            const int row = row_g + rr;
            if (total[rr] == 0) continue;
            unsigned short* cnt = s_cnt[row];
            const bool multi = total[rr] > 32;
            if (multi) {  // the counts become the number of entries already placed, per kernel point
                __syncwarp();
                if (lane < 15) cnt[lane] = 0;
                __syncwarp();
            }
            for (int e0 = 0; e0 < total[rr]; e0 += 32) {
                const int e = e0 + lane;
                const bool act = e < total[rr];
                const unsigned am = __ballot_sync(0xffffffffu, act);
                if (act) {
                    int2 rec = e0 == 0 ? first[rr] : src[rr][e];
                    const int k = (int)((unsigned)rec.x >> K_SHIFT);
                    const unsigned peers = __match_any_sync(am, k);
                    const int rank = (multi ? (int)cnt[k] : 0) + __popc(peers & lt_mask);
                    rec.x = (int)(((unsigned)rec.x & JS_MASK) | ((unsigned)row << ROW_SHIFT));
                    dst[s_start[k] + s_pos[k][row] + rank] = rec;
                    if (multi) {
                        __syncwarp(am);
                        if ((peers & lt_mask) == 0u) cnt[k] = (unsigned short)(cnt[k] + __popc(peers));
                        __syncwarp(am);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- weight pack
// images[chunk][nblk] : NB rows (output channels) x CK reduction columns, K-major core-matrix layout:
//   element (n, col) at (n/8)*128 + (col/4)*(NB*16) + (n%8)*16 + (col%4)*4 bytes.  col <-> (k, c) = (col / cin_p, col % cin_p)
// value = tf32(W[k*sk + c*sc + n*sn]) inside the valid range, else 0.
struct PackJob {
    const float* W;
    float* images;
    long long sk, sc, sn;
    int K, cin, cin_p, cout, NB, n_nblk, n_chunks;
    long long first;  // first destination float of this job in the launch-wide enumeration
};
constexpr int PACK_MAX_JOBS = 64;
struct PackJobs {
    int n;
    long long total;
    PackJob job[PACK_MAX_JOBS];
};

__device__ __forceinline__ void pack_one(const PackJob& J, long long t) {
    const long long per_img = (long long)J.NB * FWD_CK;
    const long long img = t / per_img;
    const int r = (int)(t - img * per_img);
    const int chunk = (int)(img / J.n_nblk), nblk = (int)(img - (long long)chunk * J.n_nblk);
    const int j = r / (J.NB * 4);          // 16-byte K chunk
    const int rem = r - j * (J.NB * 4);
    const int n8 = rem >> 5, in8 = rem & 31;
    const int n = n8 * 8 + (in8 >> 2), e = in8 & 3;
    const int col = chunk * FWD_CK + j * 4 + e;
    const int k = col / J.cin_p, c = col - k * J.cin_p;
    const int ng = nblk * J.NB + n;
    float v = 0.f;
    if (k < J.K && c < J.cin && ng < J.cout) v = to_tf32(J.W[k * J.sk + c * J.sc + ng * J.sn]);
    J.images[t] = v;
}

// every job of a launch in one grid-stride enumeration of destination floats (coalesced stores); the job of an element
// is found by a short linear search over the (few) job boundaries
__global__ void __launch_bounds__(256) kp_pack_w_kernel(const __grid_constant__ PackJobs P) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < P.total; t += (long long)gridDim.x * blockDim.x) {
        int j = 0;
        while (j + 1 < P.n && t >= P.job[j + 1].first) j++;
        pack_one(P.job[j], t - P.job[j].first);
    }
}

// zero-pad the channel dimension (float4 gathers need a row pitch the lane groups divide)
__global__ void __launch_bounds__(256) kp_pad_cols_kernel(const float* __restrict__ src, long long rows, int c, int c_p,
                                                         float* __restrict__ dst) {
    const long long total = rows * c_p;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / c_p;
        const int cc = (int)(t - r * c_p);
        dst[t] = cc < c ? src[r * c + cc] : 0.f;
    }
}

// ------------------------------------------------------------------------------------- A-stage producers (fwd, dW)
// Where the 16 bytes (4 consecutive reduction columns = column group cg) of stage row p live in shared memory.
struct LayoutKMajor {  // forward: UMMA K-major, no swizzle (see A_SBO / A_LBO above)
    static __device__ __forceinline__ int off(int p, int cg) { return cg * A_LBO + (p >> 3) * A_SBO + (p & 7) * 16; }
};
struct LayoutMNMajor {  // dW: M = reduction column (cg = 4 of them), K = point p (see MN_SBO / MN_LBO above)
    static __device__ __forceinline__ int off(int p, int cg) {
        return (cg >> 3) * MN_LBO + (p >> 2) * MN_SBO + (p & 3) * 128 + ((((cg & 7) >> 1) ^ (p & 3)) << 5) + (cg & 1) * 16;
    }
};

// How a stage is shared out. A stage is (rows x columns) = 128 x 64 (forward) or 64 x 128 (dW): 2048 float4 cells for
// the 256 producer lanes, i.e. every lane owns ONE column group (4 reduction columns = one kernel point, 4 channels) of
// ONE block of 8 rows. Lanes are bundled into groups of G = seg_len / 4 (seg_len = min(Cin, 64) consecutive columns of
// one kernel point): a group owns a (segment, row block) unit and walks exactly the entries of the tile's flat list of
// that kernel point that fall into the row block, toff[k][rb] .. toff[k][rb + 1]: sorted by row, so the lane accumulates
// in registers while the row stays the same and stores a finished row ONCE (rows without entries are stored as zeros on
// the way): no zero-fill pass, no read-modify-write of shared memory, no per-row pointer chasing, and the entry records
// of a unit are consecutive in memory. U entries are in flight per lane (their records first, then their float4 gathers).
struct GatherGeom {
    int g_log2;      // log2(G)
    int seg_len;     // min(cin_p, 64)
    int cin_p, K;
};

// RPL = rows per lane: 8 in every current instantiation (a stage is filled by 8 warps); 4 = two lane groups share a
// unit, each walks the unit's entries and takes the rows of its half (kept for experiments with 16 warps per stage).
template <class LAY, int NRBS, int U, int RPL>
__device__ __forceinline__ void produce_sparse(unsigned char* sA, int warp, int lane, const GatherGeom& gg, int col_base,
                                               int rb_base, const int* __restrict__ toff_tile,
                                               const int2* __restrict__ ent, const float* __restrict__ x) {
    const int G = 1 << gg.g_log2;
    const int gi = (warp << (5 - gg.g_log2)) + (lane >> gg.g_log2);
    const int g = RPL == RB ? gi : (gi >> 1);
    const int r_lo = RPL == RB ? 0 : (gi & 1) * RPL, r_hi = r_lo + RPL;
    const int li = lane & (G - 1);
    const int s = g / NRBS, rb = g - s * NRBS;
    const int cg = s * G + li;
    const int col = col_base + s * gg.seg_len;
    const int k = col / gg.cin_p;
    const int c0 = col - k * gg.cin_p + 4 * li;
    int a = 0, b = 0;
    if (k < gg.K) {
        a = __ldg(toff_tile + k * TOFF_RB + rb_base + rb);
        b = __ldg(toff_tile + k * TOFF_RB + rb_base + rb + 1);
    }
    const int2* __restrict__ ep = ent + a;
    const int n = b - a;
    const float* __restrict__ xc = x + c0;
    const int p0 = rb * RB;
    // The lane's cells are zeroed first, so that a finished row is ONE unconditional store and rows without entries need
    // no bookkeeping: the earlier version, which stored zeros for the gaps between rows on the way, spent ~200 SASS
    // instructions per entry slot on unrolled gap loops, and this loop is bound by its instruction count (the warps of an
    // SM sub-partition run it in the same phase; ncu: issue slots, not memory).
#pragma unroll
    for (int r = 0; r < RPL; r++)
        *reinterpret_cast<float4*>(sA + LAY::off(p0 + r_lo + r, cg)) = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = r_lo;
    int2 rec[U];
#pragma unroll
    for (int u = 0; u < U; u++) rec[u] = u < n ? __ldg(ep + u) : make_int2(0, 0);
    for (int e0 = 0; e0 < n; e0 += U) {
        float4 xv[U];
        float wv[U];
        int rv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            wv[u] = __int_as_float(rec[u].y);
            rv[u] = (int)(((unsigned)rec[u].x >> ROW_SHIFT) & (RB - 1));
            if (e0 + u >= n || (RPL != RB && (rv[u] < r_lo || rv[u] >= r_hi))) rv[u] = -1;   // not mine
            if (rv[u] >= 0) xv[u] = __ldg(reinterpret_cast<const float4*>(xc + (size_t)((unsigned)rec[u].x & J_MASK) * gg.cin_p));
        }
#pragma unroll
        for (int u = 0; u < U; u++) rec[u] = e0 + U + u < n ? __ldg(ep + e0 + U + u) : make_int2(0, 0);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rv[u] >= 0) {
                if (rv[u] != cur) {
                    *reinterpret_cast<float4*>(sA + LAY::off(p0 + cur, cg)) = to_tf32(acc);
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cur = rv[u];
                }
                acc.x = fmaf(wv[u], xv[u].x, acc.x); acc.y = fmaf(wv[u], xv[u].y, acc.y);
                acc.z = fmaf(wv[u], xv[u].z, acc.z); acc.w = fmaf(wv[u], xv[u].w, acc.w);
            }
        }
    }
    *reinterpret_cast<float4*>(sA + LAY::off(p0 + cur, cg)) = to_tf32(acc);
}

// Dense variant (linear layers next to KPConv: the A operand is a plain row-major matrix a[n, ld]): the lane's 8 cells
// are 8 independent float4 loads, optionally scaled by the LeakyReLU derivative taken from `mask` (same shape: factor 1
// where mask > 0, `slope` elsewhere), rounded to TF32.
template <class LAY, int NRBS, int RPL>
__device__ __forceinline__ void produce_dense(unsigned char* sA, int warp, int lane, int col_base, int row_base, int n,
                                              const float* __restrict__ a, int ld, const float* __restrict__ mask,
                                              float slope) {
    constexpr int NCG = 256 / NRBS;                 // column groups per stage (16 forward, 32 dW)
    const int t = warp * 32 + lane;
    const int cg = t % NCG;
    const int col = col_base + 4 * cg;
    const int p0 = (t / NCG) * RPL;
    float4 v[RPL];
#pragma unroll
    for (int r = 0; r < RPL; r++) {
        const int i = row_base + p0 + r;
        v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n && col < ld) v[r] = __ldg(reinterpret_cast<const float4*>(a + (size_t)i * ld + col));
    }
    if (mask) {
#pragma unroll
        for (int r = 0; r < RPL; r++) {
            const int i = row_base + p0 + r;
            if (i < n && col < ld) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(mask + (size_t)i * ld + col));
                v[r].x *= y.x > 0.f ? 1.f : slope; v[r].y *= y.y > 0.f ? 1.f : slope;
                v[r].z *= y.z > 0.f ? 1.f : slope; v[r].w *= y.w > 0.f ? 1.f : slope;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPL; r++) *reinterpret_cast<float4*>(sA + LAY::off(p0 + r, cg)) = to_tf32(v[r]);
}

// --------------------------------------------------------------------------------------------------------- forward
struct FwdParams {
    int nq;              // centre points (rows of out)
    const float* x;      // [n_other, cin_p] features gathered through the lists (dense mode: the A matrix [nq, cin_p])
    GatherGeom gg;
    const int* toff;     // tile headers (sparse mode)
    const int2* entries;
    const float* images; // packed weights [n_chunks][n_nblk][NB*64]
    int NB, n_nblk, n_chunks;
    int ksplit;          // CTAs along the reduction (blockIdx.y); > 1 => partial sums are added atomically into out
    int stages;
    int cluster;         // CTAs per cluster along x (tiles): 1, or 2 / 4 = the weight chunks are multicast inside the cluster
    int n_tiles;         // real tiles (gridDim.x is rounded up to the cluster size; the extra CTAs only take part in the copies)
    float* out;          // [nq, cout] with row stride ldo (pre-zeroed when ksplit > 1)
    int cout, ldo;
    uint32_t tmem_cols;
    // dense mode (template DENSE): optional LeakyReLU-derivative mask on the A matrix
    const float* mask;
    float slope_in;
    // epilogue (ksplit == 1 only): out = leaky(acc + bias, slope_out); bias may be null, slope_out = 1 disables
    const float* bias;
    float slope_out;
};

// shared memory: [stages] A stages, [stages] B stages, barriers. Barrier use: a_full[s] counts the 8 producer warps,
// b_full[s] the loader's expect_tx + the bulk copy's bytes; a_empty / b_empty / acc_full are arrived on by tcgen05.commit.
struct FwdBars {
    uint64_t a_full[MAX_STAGES], a_empty[MAX_STAGES], b_full[MAX_STAGES], b_empty[MAX_STAGES], acc_full;
    uint32_t tmem;
};

template <bool DENSE, int NPW>
__global__ void __launch_bounds__((NPW + 2) * 32, NPW == 8 ? 2 : 1) kp_fwd_kernel(const __grid_constant__ FwdParams P) {
    constexpr int LOADER_WARP = NPW, MMA_WARP = NPW + 1, NGRP = NPW / 8;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int S = P.stages;
    const int b_bytes = P.NB * FWD_CK * 4;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)S * A_STAGE;
    FwdBars* bars = reinterpret_cast<FwdBars*>(sB + (size_t)S * b_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, tile_base = tile * TILE_M;
    const int cps = (P.n_chunks + P.ksplit - 1) / P.ksplit;
    const int c0 = blockIdx.y * cps, c1 = min(P.n_chunks, c0 + cps);
    if (c0 >= c1) return;   // (uniform over a cluster: its CTAs share blockIdx.y)
    const int n_loc = c1 - c0;
    const int CL = P.cluster;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    const bool ghost = tile >= P.n_tiles;   // padding CTA of the last cluster: copies its weight slices, nothing else

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&bars->a_full[s], 8);
            mbar_init(&bars->a_empty[s], 1);
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], CL);   // every CTA of the cluster must have consumed a stage before it is refilled
        }
        mbar_init(&bars->acc_full, 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&bars->tmem, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // barriers of all CTAs initialised before anyone copies into / arrives on them
    tc_fence_after();
    const uint32_t tmem = bars->tmem;

    if (warp < NPW) {
        // ===== producers: A chunk `it` -> stage it % S =====
        const int* toff_tile = nullptr;
        const int2* ent = nullptr;
        if (!DENSE && !ghost) {
            toff_tile = P.toff + (size_t)tile * TOFF_PER_TILE;
            ent = P.entries + __ldg(toff_tile + 15 * TOFF_RB);
        }
        // a stage is always filled by 8 warps; with 16 producer warps (one CTA per SM) the two groups of 8 take
        // alternate chunks, so two stages are being assembled at any time
        const int pw = warp & 7;
        for (int it = warp >> 3; it < n_loc && !ghost; it += NGRP) {
            const int s = it % S, use = it / S;
            if (use > 0) mbar_wait(&bars->a_empty[s], (uint32_t)((use - 1) & 1));
            unsigned char* a = sA + (size_t)s * A_STAGE;
            const int col_base = (c0 + it) * FWD_CK;
            if (DENSE) produce_dense<LayoutKMajor, NRB, RB>(a, pw, lane, col_base, tile_base, P.nq, P.x, P.gg.cin_p, P.mask, P.slope_in);
            else produce_sparse<LayoutKMajor, NRB, KP_U, RB>(a, pw, lane, P.gg, col_base, 0, toff_tile, ent, P.x);
            fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->a_full[s]);
        }
    } else if (warp == LOADER_WARP) {
        // ===== loader: packed weight chunk of step t -> B stage t % S =====
        if (lane == 0) {
            const int n_steps = n_loc * P.n_nblk;
            for (int t = 0; t < n_steps; t++) {
                const int s = t % S, use = t / S;
                if (use > 0) mbar_wait(&bars->b_empty[s], (uint32_t)((use - 1) & 1));
                const int chunk = c0 + t / P.n_nblk, nblk = t % P.n_nblk;
                mbar_expect_tx(&bars->b_full[s], (uint32_t)b_bytes);
                const float* src = P.images + ((size_t)chunk * P.n_nblk + nblk) * (size_t)P.NB * FWD_CK;
                if (CL == 1) bulk_g2s(sB + (size_t)s * b_bytes, src, (uint32_t)b_bytes, &bars->b_full[s]);
                else {
                    const uint32_t slice = (uint32_t)b_bytes / (uint32_t)CL;
                    bulk_g2s_multicast(sB + (size_t)s * b_bytes + (size_t)crank * slice,
                                       reinterpret_cast<const unsigned char*>(src) + (size_t)crank * slice, slice,
                                       &bars->b_full[s], cmask);
                }
            }
        }
    } else {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(TILE_M, P.NB, 0, 0);
            const uint32_t b_lbo = (uint32_t)P.NB * 16u;
            int t = 0;
            if (ghost) {  // release every stage as soon as it has arrived (no MMAs outstanding: the commit arrives at once)
                for (int tt = 0; tt < n_loc * P.n_nblk; tt++) {
                    mbar_wait(&bars->b_full[tt % S], (uint32_t)((tt / S) & 1));
                    umma_commit_multicast(&bars->b_empty[tt % S], cmask);
                }
            }
            for (int it = 0; it < n_loc && !ghost; it++) {
                const int s = it % S;
                mbar_wait(&bars->a_full[s], (uint32_t)((it / S) & 1));
                const uint32_t a_addr = smem_u32(sA + (size_t)s * A_STAGE);
                for (int nblk = 0; nblk < P.n_nblk; nblk++, t++) {
                    const int sb = t % S;
                    mbar_wait(&bars->b_full[sb], (uint32_t)((t / S) & 1));
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(sB + (size_t)sb * b_bytes);
#pragma unroll
                    for (int kk = 0; kk < FWD_CK / 8; kk++) {  // K = 8 per tf32 MMA = two 16-byte K chunks
                        const uint64_t ad = make_desc(a_addr + kk * 2 * A_LBO, A_LBO, A_SBO);
                        const uint64_t bd = make_desc(b_addr + kk * 2 * b_lbo, b_lbo, B_SBO);
                        umma_tf32(tmem + (uint32_t)(nblk * P.NB), ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    if (CL == 1) umma_commit(&bars->b_empty[sb]);
                    else umma_commit_multicast(&bars->b_empty[sb], cmask);
                }
                umma_commit(&bars->a_empty[s]);
            }
            if (!ghost) umma_commit(&bars->acc_full);
        }
    }

    // ===== epilogue: the producer warps drain TMEM (warp w: lanes 32*(w%4).., column blocks of 16 dealt over w/4) =====
    if (warp < NPW && !ghost) {
        mbar_wait(&bars->acc_full, 0u);
        __syncwarp();
        tc_fence_after();
        const int row = tile_base + 32 * (warp & 3) + lane;
        const int n_cb = (P.n_nblk * P.NB) / 16;
        const bool vec = (P.cout & 3) == 0 && (P.ldo & 3) == 0;
        const bool post = P.ksplit == 1 && (P.bias != nullptr || P.slope_out != 1.f);
        for (int cb = warp >> 2; cb < n_cb; cb += NPW / 4) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cb * 16), v);
            if (post) {
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    float a = v[t];
                    if (P.bias && cb * 16 + t < P.cout) a += __ldg(P.bias + cb * 16 + t);
                    v[t] = a > 0.f ? a : a * P.slope_out;
                }
            }
            if (row < P.nq) {
                float* o = P.out + (size_t)row * P.ldo + cb * 16;
#pragma unroll
                for (int t = 0; t < 16; t += 4) {
                    if (vec && cb * 16 + t < P.cout) {
                        const float4 f = make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]);
                        if (P.ksplit > 1) atomicAdd(reinterpret_cast<float4*>(o + t), f);
                        else *reinterpret_cast<float4*>(o + t) = f;
                    } else if (!vec) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (cb * 16 + t + u < P.cout) {
                                if (P.ksplit > 1) atomicAdd(o + t + u, v[t + u]);
                                else o[t + u] = v[t + u];
                            }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still copy into its stages or arrive on its barriers
    if (warp == MMA_WARP) tmem_dealloc(tmem, P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------- dW
// B stage = dOut[64 points, NB outputs], MN-major SWIZZLE_128B_BASE32B like the A stage (m = output column).
__host__ __device__ constexpr int dw_b_bytes(int NB) { return ((NB + 31) / 32) * MN_LBO; }

struct DwParams {
    int nq;
    const float* x;      // [ns, cin_p]
    GatherGeom gg;
    int cin;
    const int* toff;
    const int2* entries;
    const float* dout;   // [nq, cout]
    int cout, NB;        // NB = output columns handled per CTA (<= 256), slice index = blockIdx.z
    int n_tiles, n_splits, stages;
    int b_tma;           // dOut stages are filled by the loader warp with tensor copies (else by the producer warps)
    float* dw;           // [K, cin, cout], pre-zeroed
    uint32_t tmem_cols;
    // dense mode: A = x[nq, cin_p] itself (times the LeakyReLU derivative read from mask), see produce_dense
    const float* mask;
    float slope_in;
};

struct DwBars {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES], acc_full;
    uint32_t tmem;
};

template <bool DENSE, int NPW>
__global__ void __launch_bounds__((NPW + 2) * 32, NPW == 8 ? 2 : 1) kp_dw_kernel(const __grid_constant__ DwParams P,
                                                                               const __grid_constant__ CUtensorMap dout_map) {
    constexpr int LOADER_WARP = NPW, MMA_WARP = NPW + 1, NGRP = NPW / 8;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int S = P.stages;
    const int b_bytes = dw_b_bytes(P.NB);
    const int stage_bytes = DW_A_STAGE + b_bytes;  // both multiples of 1024 (the swizzle uses address bits)
    DwBars* bars = reinterpret_cast<DwBars*>(smem + (size_t)S * stage_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, split = blockIdx.y, n0 = blockIdx.z * P.NB;
    if (split >= P.n_tiles) return;  // uniform: nothing to do for this CTA
    const int my_tiles = (P.n_tiles - split + P.n_splits - 1) / P.n_splits;
    const int n_steps = my_tiles * (TILE_M / DW_PT);

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&bars->full[s], P.b_tma ? 9 : 8);  // 8 producer warps (+ the loader's expect_tx arrival)
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->acc_full, 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(&bars->tmem, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem;

    if (warp < NPW) {
        // ===== producers: step t = (tile, half): A stage [64 points x 128 columns] + dOut stage [64 points x NB] =====
        const int nv = P.NB >> 2;
        const bool vec4 = (P.cout & 3) == 0;
        const int pw = warp & 7;  // (groups of 8 warps take alternate steps, as in the forward kernel)
        for (int t = warp >> 3; t < n_steps; t += NGRP) {
            const int s = t % S, use = t / S;
            const int tile = split + (t >> 1) * P.n_splits, half = t & 1;
            const int row_base = tile * TILE_M + half * DW_PT;
            if (use > 0) mbar_wait(&bars->empty[s], (uint32_t)((use - 1) & 1));
            unsigned char* a = smem + (size_t)s * stage_bytes;
            unsigned char* b = a + DW_A_STAGE;
            // dOut rows -> B stage (TF32): 64 rows x NB/4 float4 items dealt over the 256 lanes, 4 loads in flight
            // (only when the loader warp cannot do it: row pitch of dOut not a multiple of 16 bytes)
            const int items = P.b_tma ? 0 : DW_PT * nv;
            for (int base = pw * 32 + lane; base < items; base += 4 * 256) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = base + u * 256;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < items) {
                        const int r = q / nv, n4 = q - r * nv;
                        const int i = row_base + r, n = n0 + n4 * 4;
                        if (i < P.nq && n < P.cout) {
                            const float* src = P.dout + (size_t)i * P.cout + n;
                            if (vec4) v[u] = __ldg(reinterpret_cast<const float4*>(src));
                            else {
                                v[u].x = src[0];
                                if (n + 1 < P.cout) v[u].y = src[1];
                                if (n + 2 < P.cout) v[u].z = src[2];
                                if (n + 3 < P.cout) v[u].w = src[3];
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int q = base + u * 256;
                    if (q < items) {
                        const int r = q / nv, n4 = q - r * nv;
                        *reinterpret_cast<float4*>(b + LayoutMNMajor::off(r, n4)) = to_tf32(v[u]);
                    }
                }
            }
            if (DENSE) {
                produce_dense<LayoutMNMajor, DW_PT / RB, RB>(a, pw, lane, chunk * DW_CK, row_base, P.nq, P.x, P.gg.cin_p, P.mask, P.slope_in);
            } else {
                const int* toff_tile = P.toff + (size_t)tile * TOFF_PER_TILE;
                const int2* ent = P.entries + __ldg(toff_tile + 15 * TOFF_RB);
                produce_sparse<LayoutMNMajor, DW_PT / RB, KP_U, RB>(a, pw, lane, P.gg, chunk * DW_CK, half * (DW_PT / RB), toff_tile, ent, P.x);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->full[s]);
        }
    } else if (warp == LOADER_WARP) {
        // ===== loader: dOut[64 points x NB] -> B stage, one tensor copy per 32 output columns; the copy engine applies
        //       the 32-byte-atom swizzle and the fp32 -> tf32 conversion (tensor map data type TFLOAT32), rows and
        //       columns past the end of dOut arrive as zeros =====
        if (lane == 0 && P.b_tma) {
            const int n_box = (P.NB + 31) / 32;
            for (int t = 0; t < n_steps; t++) {
                const int s = t % S, use = t / S;
                const int tile = split + (t >> 1) * P.n_splits, half = t & 1;
                const int row_base = tile * TILE_M + half * DW_PT;
                if (use > 0) mbar_wait(&bars->empty[s], (uint32_t)((use - 1) & 1));
                unsigned char* b = smem + (size_t)s * stage_bytes + DW_A_STAGE;
                mbar_expect_tx(&bars->full[s], (uint32_t)(n_box * MN_LBO));
                for (int j = 0; j < n_box; j++) tma_g2s_2d(b + (size_t)j * MN_LBO, &dout_map, n0 + 32 * j, row_base, &bars->full[s]);
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(DW_CK, P.NB, 1, 1);
            for (int t = 0; t < n_steps; t++) {
                const int s = t % S;
                mbar_wait(&bars->full[s], (uint32_t)((t / S) & 1));
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + DW_A_STAGE;
#pragma unroll
                for (int kk = 0; kk < DW_PT / 8; kk++) {  // K = 8 points per MMA
                    // MN-major swizzled descriptors: "leading" offset = next 32-element group along M/N, "stride" offset =
                    // next group of 4 K values; 8 points per MMA = two K groups = 1024 B
                    const uint64_t ad = make_desc(a_addr + kk * 2 * MN_SBO, MN_LBO, MN_SBO, LAYOUT_SW128_BASE32B);
                    const uint64_t bd = make_desc(b_addr + kk * 2 * MN_SBO, MN_LBO, MN_SBO, LAYOUT_SW128_BASE32B);
                    umma_tf32(tmem, ad, bd, idesc, (t > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->acc_full);
        }
    }

    // epilogue: TMEM lane r = reduction column chunk*128 + r = (k, c); add into dW[k, c, n0 + col]
    if (warp < NPW) {
        mbar_wait(&bars->acc_full, 0u);
        __syncwarp();
        tc_fence_after();
        const int r = 32 * (warp & 3) + lane;
        const int col = chunk * DW_CK + r;
        const int k = col / P.gg.cin_p, c = col - k * P.gg.cin_p;
        const bool row_ok = k < P.gg.K && c < P.cin;
        const bool vec = (P.cout & 3) == 0;
        for (int cb = warp >> 2; cb < P.NB / 16; cb += NPW / 4) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cb * 16), v);
            if (row_ok) {
                float* o = P.dw + ((size_t)k * P.cin + c) * P.cout + n0 + cb * 16;
#pragma unroll
                for (int t = 0; t < 16; t += 4) {
                    if (vec && n0 + cb * 16 + t < P.cout) {
                        atomicAdd(reinterpret_cast<float4*>(o + t), make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]));
                    } else if (!vec) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (n0 + cb * 16 + t + u < P.cout) atomicAdd(o + t + u, v[t + u]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem, P.tmem_cols);
}

// ----------------------------------------------------------------------------------------- transposed neighbour table
template <typename IdxT>
__global__ void __launch_bounds__(256) kp_tr_count_kernel(const IdxT* __restrict__ idx, int nq, int H, int stride, int ns,
                                                         int* __restrict__ deg) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)nq * H) return;
    const long long j = (long long)idx[(t / H) * stride + (t % H)];
    if (j >= 0 && j < ns) atomicAdd(&deg[j], 1);
}

template <typename IdxT>
__global__ void __launch_bounds__(256) kp_tr_fill_kernel(const IdxT* __restrict__ idx, int nq, int H, int stride, int ns,
                                                        const int* __restrict__ rowptr, int* __restrict__ cursor,
                                                        int* __restrict__ col) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)nq * H) return;
    const long long j = (long long)idx[(t / H) * stride + (t % H)];
    if (j >= 0 && j < ns) col[rowptr[j] + atomicAdd(&cursor[j], 1)] = (int)(t / H);
}

// each row's centre list put in ascending order (rank by counting, one warp per row; the entries of a row are
// distinct), so that the dX sums run in a fixed order whatever order the atomics above filled the row in
__global__ void __launch_bounds__(256) kp_tr_sort_kernel(const int* __restrict__ rowptr, int ns,
                                                        const int* __restrict__ col, int* __restrict__ col_sorted,
                                                        int* __restrict__ rowptr_last, const int* __restrict__ total) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j == 0 && lane == 0) *rowptr_last = *total;  // rowptr[ns]
    if (j >= ns) return;
    const int a = rowptr[j];
    const int b = (j + 1 < ns) ? rowptr[j + 1] : *total;
    if (lane == 0 && b - a > *(volatile int*)(rowptr_last + 1)) atomicMax(rowptr_last + 1, b - a);  // rowptr[ns+1]: longest row
    for (int e = a + lane; e < b; e += 32) {
        const int v = col[e];
        int rank = 0;
        for (int u = a; u < b; u++) rank += (col[u] < v) ? 1 : 0;
        col_sorted[a + rank] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------- host side
static int pad4(int c) { return (c + 3) & ~3; }
// row pitch of a gathered matrix: the lane groups of the producers cover min(pitch, 64) columns of one kernel point, so
// the pitch is a power of two up to 64 or a multiple of 64
static int pad_gather(int c) {
    if (c > 64) return (c + 63) & ~63;
    int p = 4;
    while (p < c) p <<= 1;
    return p;
}
static GatherGeom make_geom(int cin_p, int K) {
    GatherGeom g;
    g.cin_p = cin_p;
    g.K = K;
    g.seg_len = cin_p < FWD_CK ? cin_p : FWD_CK;
    g.g_log2 = 0;
    while ((4 << g.g_log2) < g.seg_len) g.g_log2++;
    return g;
}

// Caller-visible list buffers: `hdr` = 4 control ints (entries used, overflow flag, 2 spare) followed by the tile
// headers, `entries` = the compact entry array.
constexpr int LISTS_CTL_INTS = 4;
struct Lists {
    int* hdr;
    int2* entries;
    long long cap;
    const int* toff() const { return hdr + LISTS_CTL_INTS; }
};

void kpconv_lists_bytes(int nc, long long n_pairs, long long* hdr_bytes, long long* entries_bytes) {
    const long long tiles = nc > 0 ? (nc + TILE_M - 1) / TILE_M : 1;
    *hdr_bytes = (LISTS_CTL_INTS + tiles * TOFF_PER_TILE) * 4;
    *entries_bytes = (n_pairs > 0 ? n_pairs : 1) * 15 * 8;  // exact worst case; callers with calibrated bounds pass less
}

// host side of a batch of list jobs: add_list_job() per job, then launch_list_jobs()
struct ListBatch {
    ListJobs J;
    int n_tiles;
    long long cost[LST_MAX_JOBS];  // table cells per tile (long rows first: they finish last)
    ListBatch() : n_tiles(0) { J.n = 0; }
};
static int launch_list_jobs(ListBatch& B, cudaStream_t stream) {
    if (B.J.n == 0) return KP_OK;
    // jobs with the longest rows go first in the grid, so the last wave is made of cheap tiles
    for (int a = 1; a < B.J.n; a++)
        for (int b = a; b > 0 && B.cost[b] > B.cost[b - 1]; b--) {
            std::swap(B.cost[b], B.cost[b - 1]);
            std::swap(B.J.job[b], B.J.job[b - 1]);
        }
    int first = 0;
    for (int a = 0; a < B.J.n; a++) {
        B.J.job[a].first_tile = first;
        first += ceil_div(B.J.job[a].nc, TILE_M);
    }
    ProfileScope ps("kp_lists", stream);
    kp_lists_ctl_kernel<<<1, 128, 0, stream>>>(B.J);
    KP_CHECK_LAUNCH();
    kp_lists_kernel<<<first, LST_WARPS * 32, 0, stream>>>(B.J);
    KP_CHECK_LAUNCH();
    B.J.n = 0;
    B.n_tiles = 0;
    return KP_OK;
}
static int add_list_job(ListBatch& B, Scratch& S, const float* centres, int nc, const float* others, int no, const Table& T,
                        long long n_pairs, const float* kp, int K, float kp_sign, float extent, Lists* L,
                        cudaStream_t stream, int* overflow = nullptr) {
    if (n_pairs * 15 >= (1LL << 31)) return fail(KP_ERR_UNSUPPORTED, "kpconv: neighbour table too large (Nq*H*15 >= 2^31)");
    const int n_tiles = ceil_div(nc, TILE_M);
    if (!L->hdr) {
        L->hdr = S.alloc<int>((size_t)LISTS_CTL_INTS + (size_t)n_tiles * TOFF_PER_TILE);
        L->cap = (n_pairs > 0 ? n_pairs : 1) * 15;
        L->entries = S.alloc<int2>((size_t)L->cap);
    }
    int2* scratch = S.alloc<int2>((size_t)(n_pairs > 0 ? n_pairs : 1) * 15);
    if (S.status != KP_OK) return S.status;
    if (B.J.n == LST_MAX_JOBS) {
        const int rc = launch_list_jobs(B, stream);
        if (rc != KP_OK) return rc;
    }
    ListJob& j = B.J.job[B.J.n];
    j.centres = centres; j.others = others; j.T = T; j.kp = kp; j.scratch = scratch;
    j.L.toff = L->hdr + LISTS_CTL_INTS; j.L.entries = L->entries; j.L.ctl = L->hdr; j.L.cap = L->cap; j.L.overflow = overflow;
    j.nc = nc; j.no = no; j.K = K; j.kp_sign = kp_sign; j.inv_ext = 1.f / extent; j.first_tile = 0;
    B.cost[B.J.n] = nc > 0 ? n_pairs / nc : 0;
    B.J.n++;
    B.n_tiles += n_tiles;
    return KP_OK;
}
static int build_lists(Scratch& S, const float* centres, int nc, const float* others, int no, const Table& T,
                       long long n_pairs, const float* kp, int K, float kp_sign, float extent, Lists* L,
                       cudaStream_t stream, int* overflow = nullptr) {
    ListBatch B;
    const int rc = add_list_job(B, S, centres, nc, others, no, T, n_pairs, kp, K, kp_sign, extent, L, stream, overflow);
    return rc != KP_OK ? rc : launch_list_jobs(B, stream);
}

// How many CTAs share the reduction axis of one 128-point tile. A CTA costs (its chunks + ~1 chunk of fixed work:
// TMEM allocation, pipeline fill, epilogue), CTAs run in waves of `slots` (2 per SM when two CTAs fit an SM, else 1),
// and a split pays a zero-fill plus an atomic epilogue. WEASAL_KSPLIT_MODEL=0: fill the slots without looking at waves.
static int pick_ksplit(int n_tiles, int n_chunks, int slots) {
    static const bool model = !(getenv("WEASAL_KSPLIT_MODEL") && atoi(getenv("WEASAL_KSPLIT_MODEL")) == 0);
    static const bool deterministic = getenv("WEASAL_KPCONV_DETERMINISTIC") && atoi(getenv("WEASAL_KPCONV_DETERMINISTIC")) != 0;
    if (deterministic) return 1;  // one CTA per tile: bit-reproducible sums, slower on the deep layers
    if (!model) {
        int ks = ceil_div(2 * 148, n_tiles);
        if (ks > n_chunks) ks = n_chunks;
        if (ks > 16) ks = 16;
        if (ks < 1) ks = 1;
        return ceil_div(n_chunks, ceil_div(n_chunks, ks));  // no empty splits
    }
    int best = 1;
    double best_cost = 1e30;
    for (int ks = 1; ks <= n_chunks && ks <= 16; ks++) {
        const int cps = ceil_div(n_chunks, ks);
        if (ceil_div(n_chunks, cps) != ks) continue;  // would leave empty splits
        const double waves = (double)ceil_div((long long)n_tiles * ks, slots);
        const double cost = waves * (cps + 1.0) + (ks > 1 ? 0.5 * waves : 0.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = ks; }
    }
    return best;
}
int plan_ksplit(int n_tiles, int n_chunks, int slots) { return pick_ksplit(n_tiles, n_chunks, slots); }

// opt-in dynamic shared memory: a kernel's limit is only ever raised. The attribute belongs to the process, not to the
// calling thread (autograd runs backward passes on its own thread), so the bookkeeping is process-wide too.
template <typename KernelT>
static cudaError_t set_smem(KernelT kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> have;  // (kernel, device) -> current limit
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t& h = have[{(const void*)kernel, dev}];
    if (bytes <= h) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) h = bytes;
    return e;
}

constexpr size_t SMEM_TWO_CTAS = 113 * 1024;  // dynamic shared memory up to which two CTAs share an SM (228 KiB / 2 - 1 KiB)

// shape of the packed weight images of one contraction [*, K*cin_p] x [K*cin_p, cout]
struct ImgShape {
    int NB, n_nblk, n_chunks;
    long long floats() const { return (long long)n_chunks * n_nblk * NB * FWD_CK; }
};
static int img_shape(int K, int cin_p, int cout, ImgShape* s) {
    const int cout_p = (cout + 15) & ~15;
    s->NB = cout_p < 256 ? cout_p : 256;
    s->n_nblk = ceil_div(cout_p, s->NB);
    if (s->n_nblk * s->NB > 512) return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 512 output channels in one pass");
    s->n_chunks = ceil_div((long long)K * cin_p, FWD_CK);
    return KP_OK;
}

static int launch_pack(const PackJobs& J, cudaStream_t stream) {
    if (J.n == 0 || J.total == 0) return KP_OK;
    const int grid = ceil_div(J.total, 256) < 2368 ? ceil_div(J.total, 256) : 2368;
    ProfileScope ps("kp_pack_w", stream);
    kp_pack_w_kernel<<<grid, 256, 0, stream>>>(J);
    KP_CHECK_LAUNCH();
    return KP_OK;
}
static void add_pack_job(PackJobs& J, const float* W, float* images, long long sk, long long sc, long long sn, int K, int cin,
                         int cin_p, int cout, const ImgShape& sh) {
    PackJob& j = J.job[J.n++];
    j.W = W; j.images = images; j.sk = sk; j.sc = sc; j.sn = sn; j.K = K; j.cin = cin; j.cin_p = cin_p; j.cout = cout;
    j.NB = sh.NB; j.n_nblk = sh.n_nblk; j.n_chunks = sh.n_chunks;
    j.first = J.total;
    J.total += sh.floats();
}

// out[nc, cout] (row stride ldo) = A · images, A = the sparse gather through the lists (or the dense matrix x[nc, cin_p])
static int launch_fwd(const char* tag, bool dense, int nc, const float* x, const GatherGeom& gg, const int* toff,
                      const int2* entries, const float* images, const ImgShape& sh, float* out, int cout, int ldo,
                      const float* mask, float slope_in, const float* bias, float slope_out, cudaStream_t stream,
                      int* ksplit_used = nullptr) {
    FwdParams P;
    P.nq = nc; P.x = x; P.gg = gg; P.toff = toff; P.entries = entries;
    P.images = images; P.NB = sh.NB; P.n_nblk = sh.n_nblk; P.n_chunks = sh.n_chunks;
    P.out = out; P.cout = cout; P.ldo = ldo;
    P.mask = mask; P.slope_in = slope_in; P.bias = bias; P.slope_out = slope_out;
    const int n_tiles = ceil_div(nc, TILE_M);
    const size_t per_stage = (size_t)A_STAGE + (size_t)sh.NB * FWD_CK * 4;
    const int stages = sh.NB == 128 ? 3 : 2;
    const size_t smem = stages * per_stage + sizeof(FwdBars) + 64;
    P.stages = stages;
    // split the reduction across CTAs when the tiles alone cannot fill the SMs. Split partial sums meet in `out`
    // through float atomics, so their order (the last bits of the result) varies from run to run;
    // WEASAL_KPCONV_DETERMINISTIC=1 keeps one CTA per tile.
    const int ksplit = pick_ksplit(n_tiles, sh.n_chunks, smem <= SMEM_TWO_CTAS ? 2 * 148 : 148);
    P.ksplit = ksplit;
    if (ksplit_used) *ksplit_used = ksplit;
    if (ksplit > 1) KP_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)cout * 4, nc, stream));
    uint32_t cols = 32;
    while ((int)cols < sh.n_nblk * sh.NB) cols <<= 1;
    P.tmem_cols = cols;
    // clusters (WEASAL_FWD_CLUSTER = 2 / 4): the CTAs of `cl` neighbouring tiles fetch every weight chunk once and
    // multicast it, for layers whose weight stream is a large part of the L2 traffic (wide layers, many tiles)
    static const int cl_env = getenv("WEASAL_FWD_CLUSTER") ? atoi(getenv("WEASAL_FWD_CLUSTER")) : 1;
    int cl = (cl_env == 2 || cl_env == 4) ? cl_env : 1;
    if (sh.NB * FWD_CK * 4 / cl < 4096 || n_tiles < 4 * cl || sh.NB < 128) cl = 1;
    P.cluster = cl;
    P.n_tiles = n_tiles;
    ProfileScope ps(tag, stream);
    const dim3 grid(ceil_div(n_tiles, cl) * cl, ksplit);
    const bool two = smem <= SMEM_TWO_CTAS;
#define KP_LAUNCH_FWD(D, W)                                                                  \
    do {                                                                                     \
        KP_CUDA(set_smem(kp_fwd_kernel<D, W>, smem));                                        \
        if (cl == 1) kp_fwd_kernel<D, W><<<grid, (W + 2) * 32, smem, stream>>>(P);           \
        else {                                                                               \
            cudaLaunchConfig_t cfg = {};                                                     \
            cfg.gridDim = grid; cfg.blockDim = dim3((W + 2) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream; \
            cudaLaunchAttribute at[1];                                                       \
            at[0].id = cudaLaunchAttributeClusterDimension;                                  \
            at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; \
            cfg.attrs = at; cfg.numAttrs = 1;                                                \
            KP_CUDA(cudaLaunchKernelEx(&cfg, kp_fwd_kernel<D, W>, P));                       \
        }                                                                                    \
    } while (0)
    if (dense) { if (two) KP_LAUNCH_FWD(true, 8); else KP_LAUNCH_FWD(true, 16); }
    else { if (two) KP_LAUNCH_FWD(false, 8); else KP_LAUNCH_FWD(false, 16); }
#undef KP_LAUNCH_FWD
    KP_CHECK_LAUNCH();
    return KP_OK;
}

// tensor map of a row-major fp32 matrix m[rows, cols] for boxes of [64 rows x 32 columns] in the dW stage layout
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
static bool dw_stage_map(const float* m, long long rows, int cols, CUtensorMap* map) {
    static const bool off = getenv("WEASAL_DW_TMA") && atoi(getenv("WEASAL_DW_TMA")) == 0;
    memset(map, 0, sizeof(*map));
    if (off || (cols & 3) != 0 || ((uintptr_t)m & 15) != 0 || rows <= 0) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)DW_PT};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(m), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// dw[K, cin, cout] (+)= A^T · dout, A as above; dw must be zeroed by the caller
static int launch_dw(const char* tag, bool dense, int nq, const float* x, const GatherGeom& gg, int cin, const int* toff,
                     const int2* entries, const float* dout, int cout, float* dw, const float* mask, float slope_in,
                     cudaStream_t stream) {
    const int cout_p = (cout + 15) & ~15;
    DwParams P;
    P.nq = nq; P.x = x; P.gg = gg; P.cin = cin; P.toff = toff; P.entries = entries;
    P.dout = dout; P.cout = cout;
    P.NB = cout_p < 256 ? cout_p : 256;
    const int n_slices = ceil_div(cout_p, P.NB);
    const int n_chunks = ceil_div((long long)gg.K * gg.cin_p, DW_CK);
    P.n_tiles = ceil_div(nq, TILE_M);
    const size_t per_stage = (size_t)DW_A_STAGE + (size_t)dw_b_bytes(P.NB);
    const int stages = P.NB == 128 ? 3 : 2;
    const size_t smem = stages * per_stage + sizeof(DwBars) + 64;
    P.stages = stages;
    int splits = (smem <= SMEM_TWO_CTAS ? 2 * 148 : 148) / (n_chunks * n_slices);  // one wave of CTAs
    if (splits < 1) splits = 1;
    if (splits > P.n_tiles) splits = P.n_tiles;
    P.n_splits = splits;
    alignas(64) CUtensorMap dout_map;
    P.b_tma = dw_stage_map(dout, nq, cout, &dout_map) ? 1 : 0;
    P.dw = dw;
    P.mask = mask; P.slope_in = slope_in;
    uint32_t cols = 32;
    while ((int)cols < P.NB) cols <<= 1;
    P.tmem_cols = cols;
    ProfileScope ps(tag, stream);
    const dim3 grid(n_chunks, splits, n_slices);
    const bool two = smem <= SMEM_TWO_CTAS;
#define KP_LAUNCH_DW(D, W)                                                      \
    do {                                                                        \
        KP_CUDA(set_smem(kp_dw_kernel<D, W>, smem));                            \
        kp_dw_kernel<D, W><<<grid, (W + 2) * 32, smem, stream>>>(P, dout_map);  \
    } while (0)
    if (dense) { if (two) KP_LAUNCH_DW(true, 8); else KP_LAUNCH_DW(true, 16); }
    else { if (two) KP_LAUNCH_DW(false, 8); else KP_LAUNCH_DW(false, 16); }
#undef KP_LAUNCH_DW
    KP_CHECK_LAUNCH();
    return KP_OK;
}

// rows padded to the pitch the consumers need
static int padded_to(Scratch& S, const float* src, long long n, int c, int c_p, const float** dst, cudaStream_t stream) {
    *dst = src;
    if (c_p == c) return KP_OK;
    float* p = S.alloc<float>((size_t)(n > 0 ? n : 1) * c_p);
    if (S.status != KP_OK) return S.status;
    const long long tot = n * c_p;
    if (tot > 0) {
        kp_pad_cols_kernel<<<ceil_div(tot, 256) < 2048 ? ceil_div(tot, 256) : 2048, 256, 0, stream>>>(src, n, c, c_p, p);
        KP_CHECK_LAUNCH();
    }
    *dst = p;
    return KP_OK;
}

// out[nc, cout] = (sparse gather of x[n_x_rows, cin] through L) contracted with W (strides sk, sc, sn over (k, c_in,
// n_out)), or with ready-made images when `images` is non-null
static int apply_lists(const char* tag, Scratch& S, int nc, const float* x, int n_x_rows, int cin, const Lists& L,
                       const float* W, long long sk, long long sc, long long sn, const float* images, int cout, int K,
                       float* out, float slope_out, cudaStream_t stream) {
    const int cin_p = pad_gather(cin);
    const float* xg;
    int rc = padded_to(S, x, n_x_rows, cin, cin_p, &xg, stream);
    if (rc != KP_OK) return rc;
    ImgShape sh;
    if ((rc = img_shape(K, cin_p, cout, &sh)) != KP_OK) return rc;
    if (!images) {
        float* img = S.alloc<float>((size_t)sh.floats());
        if (S.status != KP_OK) return S.status;
        PackJobs J;
        J.n = 0; J.total = 0;
        add_pack_job(J, W, img, sk, sc, sn, K, cin, cin_p, cout, sh);
        if ((rc = launch_pack(J, stream)) != KP_OK) return rc;
        images = img;
    }
    return launch_fwd(tag, false, nc, xg, make_geom(cin_p, K), L.toff(), L.entries, images, sh, out, cout, cout, nullptr,
                      1.f, nullptr, slope_out, stream);
}

static int check_args(int nq, int ns, int H, int idx_stride, int cin, int cout, int K, float extent) {
    if (nq < 0 || ns < 0 || H < 0 || idx_stride < H || cin <= 0 || cout <= 0) return fail(KP_ERR_ARG, "kpconv: bad sizes");
    if (K <= 0 || K > 15) return fail(KP_ERR_UNSUPPORTED, "kpconv: kernel_size must be 1..15");
    if (!(extent > 0.f)) return fail(KP_ERR_ARG, "kpconv: KP_extent must be positive");
    if (H > 4096) return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 4096 neighbour columns");
    if ((long long)ns >= (1LL << ROW_SHIFT) || (long long)nq >= (1LL << ROW_SHIFT))
        return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 2^25 points in one call");
    return KP_OK;
}

// Standalone list construction (the prefetch stage of a training step, or a caller that keeps the lists of a static
// geometry): centres / others are the roles of the pass the lists are for (forward and dW: centres = queries, others =
// supports, kp_sign = +1; dX: centres = supports, others = queries, table = the transposed one, kp_sign = -1).
int kpconv_lists_build_device(const float* centres, int nc, const float* others, int no, const void* idx, int idx_is_i64,
                              int H, int idx_stride, const int* rowptr, const int* col, long long n_pairs,
                              const float* kp, int K, float kp_sign, float extent, void* hdr, void* entries,
                              long long entries_cap, cudaStream_t stream) {
    if (nc <= 0 || no <= 0 || K <= 0 || K > 15 || !(extent > 0.f) || !hdr || !entries || entries_cap <= 0)
        return fail(KP_ERR_ARG, "kpconv_lists_build: bad arguments");
    if ((long long)nc >= (1LL << ROW_SHIFT) || (long long)no >= (1LL << ROW_SHIFT))
        return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 2^25 points in one call");
    Scratch S(stream);
    Table T;
    if (rowptr) { T.idx = col; T.rowptr = rowptr; T.H = 0; T.stride = 0; T.is_i64 = 0; }
    else {
        if (H <= 0 || idx_stride < H) return fail(KP_ERR_ARG, "kpconv_lists_build: bad table");
        T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
        n_pairs = (long long)nc * H;
    }
    Lists L;
    L.hdr = (int*)hdr; L.entries = (int2*)entries; L.cap = entries_cap;
    return build_lists(S, centres, nc, others, no, T, n_pairs, kp, K, kp_sign, extent, &L, stream);
}

int transpose_table_device(Scratch& S, const void* idx, int idx_is_i64, int nq, int H, int idx_stride, int ns,
                           int* rowptr, int* col_sorted, cudaStream_t stream);

// Everything the KPConv calls of one batch need besides features and weights, in ONE call (a training loop issues it
// from its prefetch thread right after the pyramid): jobs run in order on `stream`.
int kpconv_prepare_device(const kp_list_job* jobs, int n_jobs, int* overflow_flag, cudaStream_t stream) {
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail(KP_ERR_ARG, "kpconv_prepare: bad arguments");
    Scratch S(stream);
    ArenaHold hold(S);  // every job's scratch stays valid until the call returns (jobs overlap on the stream)
    if (overflow_flag) KP_CUDA(cudaMemsetAsync(overflow_flag, 0, sizeof(int), stream));
    // the transposed tables first (the lists over them need them), then every list of the batch in one launch
    for (int i = 0; i < n_jobs; i++) {
        const kp_list_job& J = jobs[i];
        if (J.kind != 1) continue;
        if (J.nc < 0 || J.no <= 0 || J.H <= 0 || J.idx_stride < J.H || !J.rowptr || !J.col)
            return fail(KP_ERR_ARG, "kpconv_prepare: bad transpose job");
        const int rc = transpose_table_device(S, J.neighb_inds, J.idx_is_i64, J.nc, J.H, J.idx_stride, J.no, J.rowptr, J.col, stream);
        if (rc != KP_OK) return rc;
    }
    ListBatch B;
    for (int i = 0; i < n_jobs; i++) {
        const kp_list_job& J = jobs[i];
        if (J.kind == 1) continue;
        if (J.kind != 0 && J.kind != 2) return fail(KP_ERR_ARG, "kpconv_prepare: unknown job kind");
        if (J.nc <= 0 || J.no <= 0 || J.K <= 0 || J.K > 15 || !(J.KP_extent > 0.f) || !J.hdr || !J.entries || J.entries_cap <= 0)
            return fail(KP_ERR_ARG, "kpconv_prepare: bad list job");
        Table T;
        long long n_pairs;
        if (J.kind == 2) {
            T.idx = J.col; T.rowptr = J.rowptr; T.H = 0; T.stride = 0; T.is_i64 = 0;
            n_pairs = J.n_pairs;
        } else {
            if (J.H <= 0 || J.idx_stride < J.H) return fail(KP_ERR_ARG, "kpconv_prepare: bad table");
            T.idx = J.neighb_inds; T.rowptr = nullptr; T.H = J.H; T.stride = J.idx_stride; T.is_i64 = J.idx_is_i64;
            n_pairs = (long long)J.nc * J.H;
        }
        Lists L;
        L.hdr = (int*)J.hdr; L.entries = (int2*)J.entries; L.cap = J.entries_cap;
        const int rc = add_list_job(B, S, J.centres, J.nc, J.others, J.no, T, n_pairs, J.kernel_points, J.K, J.kp_sign,
                                    J.KP_extent, &L, stream, overflow_flag);
        if (rc != KP_OK) return rc;
    }
    return launch_list_jobs(B, stream);
}

// lists_hdr / lists_entries (optional, caller-owned device buffers of kpconv_lists_bytes): the forward pass leaves
// the influence lists there so that backward can reuse them instead of rebuilding them.
int kpconv_forward_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                          int idx_stride, const float* x, int cin, const float* w, int cout, const float* kp, int K,
                          float extent, float* out, void* lists_hdr, void* lists_entries, cudaStream_t stream) {
    int rc = check_args(nq, ns, H, idx_stride, cin, cout, K, extent);
    if (rc != KP_OK) return rc;
    if (nq == 0) return KP_OK;
    if (ns == 0 || H == 0) {
        KP_CUDA(cudaMemsetAsync(out, 0, (size_t)nq * cout * sizeof(float), stream));
        return KP_OK;
    }
    Scratch S(stream);
    Table T;
    T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
    Lists L;
    L.hdr = (int*)lists_hdr; L.entries = (int2*)lists_entries; L.cap = (long long)nq * H * 15;
    if (!lists_hdr || !lists_entries) L.hdr = nullptr;
    rc = build_lists(S, q, nq, s, ns, T, (long long)nq * H, kp, K, 1.f, extent, &L, stream);
    if (rc != KP_OK) return rc;
    return apply_lists("kp_fwd", S, nq, x, ns, cin, L, w, (long long)cin * cout, cout, 1, nullptr, cout, K, out, 1.f, stream);
}

// Forward-type pass over ready-made lists (also the dX pass: x = dOut, lists = the transposed ones, weights read
// transposed). w_packed != 0: `w` holds ready-made images (kp_pack_weights_dev); transpose_w: W'[k][o][c] = W[k][c][o].
int kpconv_apply_lists_device(int nc, const float* x, int n_x_rows, int cin, const float* w, int w_packed, int transpose_w,
                              int cout, int K, const void* hdr, const void* entries, float* out, float slope_out,
                              cudaStream_t stream) {
    if (nc < 0 || n_x_rows < 0 || cin <= 0 || cout <= 0 || K <= 0 || K > 15 || !hdr || !entries)
        return fail(KP_ERR_ARG, "kpconv_apply_lists: bad arguments");
    if (nc == 0) return KP_OK;
    Scratch S(stream);
    Lists L;
    L.hdr = (int*)hdr; L.entries = (int2*)entries; L.cap = 0;
    // B(col = (k, c), n) with c over this call's `cin` gathered channels and n over its `cout` produced channels.
    // Forward: W[k][c][n], strides (cin*cout, cout, 1). dX (this call's cin = the conv's out_channels Co, cout = its
    // in_channels Ci): B((k, o), c) = W[k][c][o] = W[k*Ci*Co + c*Co + o], strides (Ci*Co, 1, Co) = (cin*cout, 1, cin).
    const long long sk = (long long)cin * cout;
    const long long sc = transpose_w ? 1 : cout;
    const long long sn = transpose_w ? cin : 1;
    return apply_lists(transpose_w ? "kp_fwd_dx" : "kp_fwd", S, nc, x, n_x_rows, cin, L, w_packed ? nullptr : w, sk, sc, sn,
                       w_packed ? w : nullptr, cout, K, out, slope_out, stream);
}

// dW over ready-made (query-centred) lists; d_w [K, cin, cout] is overwritten
int kpconv_dw_lists_device(int nq, const float* x, int ns, int cin, const float* dout, int cout, int K, const void* hdr,
                           const void* entries, float* dw, cudaStream_t stream) {
    if (nq < 0 || ns < 0 || cin <= 0 || cout <= 0 || K <= 0 || K > 15 || !hdr || !entries)
        return fail(KP_ERR_ARG, "kpconv_dw_lists: bad arguments");
    KP_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * cin * cout * sizeof(float), stream));
    if (nq == 0 || ns == 0) return KP_OK;
    Scratch S(stream);
    const int cin_p = pad_gather(cin);
    const float* xg;
    int rc = padded_to(S, x, ns, cin, cin_p, &xg, stream);
    if (rc != KP_OK) return rc;
    return launch_dw("kp_dw", false, nq, xg, make_geom(cin_p, K), cin, (const int*)hdr + LISTS_CTL_INTS, (const int2*)entries,
                     dout, cout, dw, nullptr, 1.f, stream);
}

// Transposed neighbour table (CSR over the supports): row j lists, in ascending order, the centres i whose row contains
// j. Depends on the index matrix only, so callers may build it once per table and hand it to every backward pass that
// uses the table. rowptr has ns + 2 entries (ns + 1 row pointers, then the longest row's length), col has nq * H.
int transpose_table_device(Scratch& S, const void* idx, int idx_is_i64, int nq, int H, int idx_stride, int ns,
                           int* rowptr, int* col_sorted, cudaStream_t stream) {
    const long long n_pairs = (long long)nq * H;
    int* deg = S.alloc<int>((size_t)2 * ns + 1);  // degrees, then the fill cursors
    int* cursor = deg + ns;
    int* total = S.alloc<int>(1);
    int* scan_tmp = S.alloc<int>(scan_tmp_ints(ns));
    int* col = S.alloc<int>((size_t)(n_pairs > 0 ? n_pairs : 1));
    if (S.status != KP_OK) return S.status;
    ProfileScope pst("kp_transpose", stream);
    KP_CUDA(cudaMemsetAsync(deg, 0, ((size_t)2 * ns + 1) * sizeof(int), stream));
    KP_CUDA(cudaMemsetAsync(rowptr + ns + 1, 0, sizeof(int), stream));
    const int grid = ceil_div(n_pairs > 0 ? n_pairs : 1, 256);
    if (idx_is_i64) kp_tr_count_kernel<long long><<<grid, 256, 0, stream>>>((const long long*)idx, nq, H, idx_stride, ns, deg);
    else kp_tr_count_kernel<int><<<grid, 256, 0, stream>>>((const int*)idx, nq, H, idx_stride, ns, deg);
    KP_CHECK_LAUNCH();
    int rc = exclusive_scan(deg, rowptr, ns, total, scan_tmp, stream);
    if (rc != KP_OK) return rc;
    if (idx_is_i64) kp_tr_fill_kernel<long long><<<grid, 256, 0, stream>>>((const long long*)idx, nq, H, idx_stride, ns, rowptr, cursor, col);
    else kp_tr_fill_kernel<int><<<grid, 256, 0, stream>>>((const int*)idx, nq, H, idx_stride, ns, rowptr, cursor, col);
    KP_CHECK_LAUNCH();
    kp_tr_sort_kernel<<<ceil_div(ns, 8), 256, 0, stream>>>(rowptr, ns, col, col_sorted, rowptr + ns, total);
    KP_CHECK_LAUNCH();
    return KP_OK;
}

int transpose_table_entry(const void* idx, int idx_is_i64, int nq, int H, int idx_stride, int ns, int* rowptr,
                          int* col_sorted, cudaStream_t stream) {
    if (nq < 0 || ns <= 0 || H < 0 || idx_stride < H) return fail(KP_ERR_ARG, "transpose_table: bad sizes");
    Scratch S(stream);
    return transpose_table_device(S, idx, idx_is_i64, nq, H, idx_stride, ns, rowptr, col_sorted, stream);
}

int kpconv_backward_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                           int idx_stride, const float* x, int cin, const float* w, int cout, const float* kp, int K,
                           float extent, const float* dout, float* dx, float* dw, const void* lists_hdr,
                           const void* lists_entries, const int* t_rowptr, const int* t_col, int table_symmetric,
                           cudaStream_t stream) {
    int rc = check_args(nq, ns, H, idx_stride, cin, cout, K, extent);
    if (rc != KP_OK) return rc;
    if (table_symmetric && nq != ns) return fail(KP_ERR_ARG, "kpconv: a symmetric table needs nq == ns");
    KP_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * cin * cout * sizeof(float), stream));
    if (nq == 0 || ns == 0 || H == 0) {
        if (ns > 0) KP_CUDA(cudaMemsetAsync(dx, 0, (size_t)ns * cin * sizeof(float), stream));
        return KP_OK;
    }
    Scratch S(stream);
    const long long n_pairs = (long long)nq * H;

    // ---- dW: lists centred on the queries (same as forward)
    {
        Table T;
        T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
        Lists L;
        L.hdr = nullptr;
        if (lists_hdr && lists_entries) {  // left behind by the forward pass
            L.hdr = (int*)lists_hdr; L.entries = (int2*)lists_entries; L.cap = n_pairs * 15;
        } else {
            rc = build_lists(S, q, nq, s, ns, T, n_pairs, kp, K, 1.f, extent, &L, stream);
            if (rc != KP_OK) return rc;
        }
        const int cin_p = pad_gather(cin);
        const float* xg;
        if ((rc = padded_to(S, x, ns, cin, cin_p, &xg, stream)) != KP_OK) return rc;
        rc = launch_dw("kp_dw", false, nq, xg, make_geom(cin_p, K), cin, L.toff(), L.entries, dout, cout, dw, nullptr, 1.f, stream);
        if (rc != KP_OK) return rc;
    }

    // ---- dX: the forward kernel on the transposed table, with W^T and -kp
    Table T;
    long long t_pairs = n_pairs;
    if (table_symmetric) {
        // queries == supports and no row was cropped: j is in row i exactly when i is in row j (the f32 distance is
        // exactly symmetric), so the table is its own transpose and no CSR copy is needed
        T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
    } else {
        const int* rowptr = t_rowptr;
        const int* col_sorted = t_col;
        if (!rowptr || !col_sorted) {
            int* rp = S.alloc<int>(ns + 2);
            int* cs = S.alloc<int>((size_t)n_pairs);
            if (S.status != KP_OK) return S.status;
            rc = transpose_table_device(S, idx, idx_is_i64, nq, H, idx_stride, ns, rp, cs, stream);
            if (rc != KP_OK) return rc;
            rowptr = rp;
            col_sorted = cs;
        }
        T.idx = col_sorted; T.rowptr = rowptr; T.H = 0; T.stride = 0; T.is_i64 = 0;
    }
    Lists L;
    L.hdr = nullptr;
    rc = build_lists(S, s, ns, q, nq, T, t_pairs, kp, K, -1.f, extent, &L, stream);
    if (rc != KP_OK) return rc;
    // W'[k][c' = o][n' = c] = W[k][c][o]
    return apply_lists("kp_fwd_dx", S, ns, dout, nq, cout, L, w, (long long)cin * cout, 1, cout, nullptr, cin, K, dx, 1.f, stream);
}

// ------------------------------------------------------------------------------------------------ weight packing API
// One launch packs the operand images of any number of contractions (a training step calls it once, right after the
// optimizer step / before the forward pass, for every KPConv (W and W^T) and every unary block).
//   kind 0: KPConv forward   W[K, cin, cout]           -> images of [K*pad(cin)] x cout
//   kind 1: KPConv dX        W[K, cin, cout] transposed -> images of [K*pad(cout)] x cin
//   kind 2: linear forward   W[cout, cin] (nn.Linear)   -> images of [pad4(cin)] x cout         (y = x W^T)
//   kind 3: linear dX        W[cout, cin]               -> images of [pad4(cout)] x cin         (dx = g W)
long long pack_image_floats(int kind, int K, int cin, int cout) {
    ImgShape sh;
    int rc;
    switch (kind) {
        case 0: rc = img_shape(K, pad_gather(cin), cout, &sh); break;
        case 1: rc = img_shape(K, pad_gather(cout), cin, &sh); break;
        case 2: rc = img_shape(1, pad4(cin), cout < 512 ? cout : 512, &sh); break;
        case 3: rc = img_shape(1, pad4(cout), cin < 512 ? cin : 512, &sh); break;
        default: return -1;
    }
    if (rc != KP_OK) return -1;
    if (kind == 2) return sh.floats() * ceil_div(cout, 512);
    if (kind == 3) return sh.floats() * ceil_div(cin, 512);
    return sh.floats();
}

int pack_weights_device(int n_jobs, const int* kinds, const float* const* weights, const int* Ks, const int* cins,
                        const int* couts, float* const* images, cudaStream_t stream) {
    PackJobs J;
    J.n = 0; J.total = 0;
    for (int i = 0; i < n_jobs; i++) {
        const int K = Ks[i], cin = cins[i], cout = couts[i];
        ImgShape sh;
        int rc;
        if (kinds[i] == 0) {
            if ((rc = img_shape(K, pad_gather(cin), cout, &sh)) != KP_OK) return rc;
            add_pack_job(J, weights[i], images[i], (long long)cin * cout, cout, 1, K, cin, pad_gather(cin), cout, sh);
        } else if (kinds[i] == 1) {
            if ((rc = img_shape(K, pad_gather(cout), cin, &sh)) != KP_OK) return rc;
            add_pack_job(J, weights[i], images[i], (long long)cin * cout, 1, cout, K, cout, pad_gather(cout), cin, sh);
        } else if (kinds[i] == 2 || kinds[i] == 3) {
            // dense layers: output slices of up to 512 columns (TMEM), one job per slice, images back to back
            const int a_cols = kinds[i] == 2 ? cin : cout, o_cols = kinds[i] == 2 ? cout : cin;
            long long off = 0;
            for (int o0 = 0; o0 < o_cols; o0 += 512) {
                const int co = o_cols - o0 < 512 ? o_cols - o0 : 512;
                if ((rc = img_shape(1, pad4(a_cols), co, &sh)) != KP_OK) return rc;
                if (J.n >= PACK_MAX_JOBS) {
                    if ((rc = launch_pack(J, stream)) != KP_OK) return rc;
                    J.n = 0; J.total = 0;
                }
                // kind 2: B(c, o) = W[o*cin + c]: sc = 1, sn = cin; kind 3: B(o, c) = W[o*cin + c]: sc = cin, sn = 1
                if (kinds[i] == 2) add_pack_job(J, weights[i] + (long long)o0 * cin, images[i] + off, 0, 1, cin, 1, a_cols, pad4(a_cols), co, sh);
                else add_pack_job(J, weights[i] + o0, images[i] + off, 0, cin, 1, 1, a_cols, pad4(a_cols), co, sh);
                off += sh.floats();
            }
            continue;
        } else return fail(KP_ERR_ARG, "pack_weights: unknown kind");
        if (J.n >= PACK_MAX_JOBS - 1) {
            if ((rc = launch_pack(J, stream)) != KP_OK) return rc;
            J.n = 0; J.total = 0;
        }
    }
    return launch_pack(J, stream);
}

// ------------------------------------------------------------------------------------------ dense linear layers
// The unary blocks around every KPConv (models/blocks.py:467-507 UnaryBlock: Linear without bias -> BatchNorm, which
// is the identity on these 2-D features or a bias when use_bn is off, blocks.py:453-465 -> LeakyReLU(0.1)) run on the same
// tcgen05 pipeline with a dense A stage: one kernel for y = leaky(x W^T + b), one for dx = (dy * leaky'(y)) W, one for
// dW = (dy * leaky'(y))^T x.
__global__ void __launch_bounds__(256) kp_bias_act_kernel(float* __restrict__ y, long long n, int cout, int ld,
                                                         const float* __restrict__ bias, float slope) {
    const long long total = n * cout;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cout;
        const int c = (int)(t - r * cout);
        float a = y[r * ld + c] + (bias ? bias[c] : 0.f);
        y[r * ld + c] = a > 0.f ? a : a * slope;
    }
}

// out[n, cout] (row stride ldo) = leaky( (a ⊙ leaky'(mask))[n, acols] · B + bias ), B(c, o) = W[c*sc + o*sn], or ready-made
// images (slices of 512 output columns back to back, pack kinds 2 / 3)
static int run_dense(const char* tag, Scratch& S, int n, const float* a, int acols, int acols_valid, const float* mask,
                     float slope_in, const float* W, long long sc, long long sn, const float* images, int cout,
                     const float* bias, float slope_out, float* out, int ldo, cudaStream_t stream) {
    long long img_off = 0;
    for (int o0 = 0; o0 < cout; o0 += 512) {  // TMEM holds 512 accumulator columns
        const int co = cout - o0 < 512 ? cout - o0 : 512;
        ImgShape sh;
        int rc = img_shape(1, acols, co, &sh);
        if (rc != KP_OK) return rc;
        const float* img = images ? images + img_off : nullptr;
        img_off += sh.floats();
        if (!img) {
            float* p = S.alloc<float>((size_t)sh.floats());
            if (S.status != KP_OK) return S.status;
            PackJobs J;
            J.n = 0; J.total = 0;
            add_pack_job(J, W + (long long)o0 * sn, p, 0, sc, sn, 1, acols_valid, acols, co, sh);
            if ((rc = launch_pack(J, stream)) != KP_OK) return rc;
            img = p;
        }
        GatherGeom gg = make_geom(acols, 1);
        int ksplit = 1;  // (launch_fwd decides the split; a split finishes bias / activation in a separate pass)
        rc = launch_fwd(tag, true, n, a, gg, nullptr, nullptr, img, sh, out + o0, co, ldo, mask, slope_in,
                        bias ? bias + o0 : nullptr, slope_out, stream, &ksplit);
        if (rc != KP_OK) return rc;
        if (ksplit > 1 && (bias || slope_out != 1.f)) {  // the split partial sums met through atomics: finish separately
            const long long total = (long long)n * co;
            const int grid = ceil_div(total, 256) < 2368 ? ceil_div(total, 256) : 2368;
            kp_bias_act_kernel<<<grid, 256, 0, stream>>>(out + o0, n, co, ldo, bias ? bias + o0 : nullptr, slope_out);
            KP_CHECK_LAUNCH();
        }
    }
    return KP_OK;
}

static int check_linear(int n, int cin, int cout, float slope) {
    if (n < 0 || cin <= 0 || cout <= 0) return fail(KP_ERR_ARG, "linear: bad sizes");
    if (!(slope == slope)) return fail(KP_ERR_ARG, "linear: bad slope");
    return KP_OK;
}

// y[n, cout] = leaky(x[n, cin] · w[cout, cin]^T + bias, slope); bias may be null; slope = 1: no activation.
// w_packed: `w` holds images made by pack kind 2.
int linear_forward_device(const float* x, int n, int cin, const float* w, int w_packed, const float* bias, int cout,
                          float slope, float* y, cudaStream_t stream) {
    int rc = check_linear(n, cin, cout, slope);
    if (rc != KP_OK || n == 0) return rc;
    Scratch S(stream);
    const float* xa;
    const int cin_p = pad4(cin);
    if ((rc = padded_to(S, x, n, cin, cin_p, &xa, stream)) != KP_OK) return rc;
    return run_dense("lin_fwd", S, n, xa, cin_p, cin, nullptr, 1.f, w_packed ? nullptr : w, 1, cin, w_packed ? w : nullptr, cout,
                     bias, slope, y, cout, stream);
}

// dx[n, cin] = g · w, dw[cout, cin] = g^T · x with g = dy ⊙ leaky'(y) (y = the forward OUTPUT, or null when the layer
// has no activation); dx may be null. dw is overwritten. w_packed: `w` holds images made by pack kind 3 (dx only).
int linear_backward_device(const float* x, int n, int cin, const float* w, int w_packed, int cout, const float* y,
                           float slope, const float* dy, float* dx, float* dw, cudaStream_t stream) {
    int rc = check_linear(n, cin, cout, slope);
    if (rc != KP_OK) return rc;
    if (dw) KP_CUDA(cudaMemsetAsync(dw, 0, (size_t)cin * cout * sizeof(float), stream));
    if (n == 0) return KP_OK;
    Scratch S(stream);
    const float *ga, *ya = y;
    const int cout_p = pad4(cout);
    if ((rc = padded_to(S, dy, n, cout, cout_p, &ga, stream)) != KP_OK) return rc;
    if (y && (rc = padded_to(S, y, n, cout, cout_p, &ya, stream)) != KP_OK) return rc;
    if (dx) {
        rc = run_dense("lin_dx", S, n, ga, cout_p, cout, ya, slope, w_packed ? nullptr : w, cin, 1, w_packed ? w : nullptr, cin,
                       nullptr, 1.f, dx, cin, stream);
        if (rc != KP_OK) return rc;
    }
    if (!dw) return KP_OK;
    // dW[o, c] = sum_i g[i, o] x[i, c]: kp_dw with A = g (M = o), B stage = x rows (N = c)
    GatherGeom gg = make_geom(cout_p, 1);
    return launch_dw("lin_dw", true, n, ga, gg, cout, nullptr, nullptr, x, cin, dw, ya, slope, stream);
}

}  // namespace kp
