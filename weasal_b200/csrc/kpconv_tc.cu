// KPConv forward / backward for sm_100a: sparse kernel-point gather on CUDA cores feeding tcgen05 (TF32 in,
// FP32 accumulate in TMEM) — replaces models/blocks.py:238-374 (rigid, 'linear' influence, 'sum' aggregation) and the
// autograd backward of that expression.
//
//   out[i,:] = sum_k ( sum_h w[i,k,h] * x[idx[i,h],:] ) @ W[k],   w = max(0, 1 - ||(s[idx[i,h]] - q[i]) - kp[k]|| / ext)
//
// Facts the design rests on (measured on ALS spheres):
//   * w is ~93 % zeros: a neighbour lies within KP_extent of ~1.03 kernel points (max 3). So the first contraction
//     is done sparsely in fp32 on CUDA cores (exact, ~15x fewer FMAs than the dense [K x H] x [H x Cin] product),
//     and only the dense second contraction [P x (K*Cin)] x [(K*Cin) x Cout] goes to the tensor cores.
//   * bf16 operands give ~1.6e-3 relative error on that contraction, above the 1e-3 parity bar; TF32 operands
//     rounded to nearest give ~4e-4. Hence kind::tf32.
//
// Kernels
//   kp_influence   one warp per centre point: influence weights of every (neighbour, kernel point) pair, compacted
//                  into per-point entry lists grouped by kernel point: entry = (neighbour index | k << 27, weight).
//                  A centre's list lives in its own slot of 15 entries per table column, so no allocation / atomics.
//   kp_pack_w      W[k,c,o] -> TF32-rounded B-operand images, one per 128-column chunk of the (k,c) reduction axis,
//                  already in the UMMA K-major core-matrix layout (so a CTA fetches a chunk with one bulk copy).
//   kp_fwd         one CTA per 128 points. Per chunk: warps assemble the A tile [128 x 128] in shared memory from the
//                  entry lists (gathered float4 rows of x, all loads of a batch in flight together), one thread issues
//                  16 tcgen05.mma (M128 x N x K8) accumulating into TMEM; epilogue tcgen05.ld -> global.
//                  Backward-dX is the same kernel run on the transposed neighbour table with W^T and -kp
//                  (atomics-free segmented scatter).
//   kp_dw          dW[(k,c),o] = sum_i WF[i,(k,c)] * dOut[i,o]: the same A tile consumed MN-major (M = (k,c) rows,
//                  K = points) against the dOut tile, accumulated in TMEM across a CTA's point tiles, then added to dW.
#include "common.cuh"

#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

namespace kp {

// ------------------------------------------------------------------------------------------------------- constants
constexpr int TILE_M = 128;      // points per CTA tile (= UMMA M)
constexpr int CK = 128;          // reduction columns per chunk
// Warps per CTA (template parameter NW of the kernels): the A-tile assembly is latency bound and wants as many warps in
// flight as the SM holds. Shapes whose shared memory lets two CTAs share an SM run 8 warps per CTA (the second CTA's
// assembly overlaps the first one's MMA / barrier phases); shapes with one CTA per SM run 16 warps.
constexpr int SMEM_TWO_CTAS = 110 * 1024;
constexpr int KOFF = 16;         // per-point cumulative entry counts per kernel point (K <= 15)
constexpr int K_SHIFT = 27;      // entry.x = neighbour index | (kernel point << 27)
constexpr unsigned J_MASK = (1u << K_SHIFT) - 1u;
// A tile, UMMA canonical no-swizzle layout: element (row p, col c) at
//   (p/8)*A_SBO + (c/4)*A_LBO + (p%8)*16 + (c%4)*4     (8 rows x 16 bytes core matrices)
// A_LBO carries 16 bytes of padding so that a warp writing one row (32 lanes x 16 B) is bank-conflict free.
constexpr int A_SBO = 128;
constexpr int A_LBO = 16 * 128 + 16;          // 2064
constexpr int A_BYTES = (CK / 4) * A_LBO;     // 66048
constexpr int B_SBO = 128;

// --------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a barrier that never completes (a malformed descriptor, a lost bulk copy) traps instead of
    // hanging the GPU; 2^26 polls is seconds, far beyond any legitimate wait here
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
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

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1):
// [0,14) start>>4, [16,30) leading (K-direction) byte offset>>4, [32,46) stride (M/N-direction) byte offset>>4
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

// Entry lists. Centre i owns the slot entries[15 * row0(i) ...), row0 = i*H (padded table) or rowptr[i] (CSR): 15
// entries per table column is the exact worst case, so the list never overflows and needs no allocator. Inside the
// slot the entries are packed, grouped by kernel point (koff[i][k] = first entry of kernel point k, koff[i][15] =
// count), each group in table-column order.
constexpr int INF_WARPS = 8;
constexpr int INF_MAX_ROW = 128;  // real neighbours staged per centre in shared memory (longer rows: two-pass path)

// any row length: lanes = neighbours, one pass to count per kernel point, one to write
__device__ __forceinline__ void influence_long_row(const float* __restrict__ others, int no, const Table& T, size_t pos0,
                                               int cnt_row, float cx, float cy, float cz, const float* s_kp,
                                               float inv_ext, unsigned short* __restrict__ koff_row,
                                               int2* __restrict__ my_entries, int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int cnt[15];
#pragma unroll
    for (int k = 0; k < 15; k++) cnt[k] = 0;
    for (int hb = 0; hb < cnt_row; hb += 32) {
        const int h = hb + lane;
        long long j = (h < cnt_row) ? table_get(T, pos0 + h) : -1;
        const bool valid = j >= 0 && j < no;
        if (!__any_sync(0xffffffffu, valid)) continue;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (valid) { rx = others[3 * j] - cx; ry = others[3 * j + 1] - cy; rz = others[3 * j + 2] - cz; }
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const float w = valid ? influence_w(rx, ry, rz, s_kp[3 * k], s_kp[3 * k + 1], s_kp[3 * k + 2], inv_ext) : 0.f;
            cnt[k] += __popc(__ballot_sync(0xffffffffu, w > 0.f));
        }
    }
    int run[16];
    int total = 0;
#pragma unroll
    for (int k = 0; k < 15; k++) { run[k] = total; total += cnt[k]; }
    run[15] = total;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) mine = (lane == k) ? run[k] : mine;
    if (lane < 16) koff_row[lane] = (unsigned short)mine;
    if (total == 0) return;
    for (int hb = 0; hb < cnt_row; hb += 32) {
        const int h = hb + lane;
        long long j = (h < cnt_row) ? table_get(T, pos0 + h) : -1;
        const bool valid = j >= 0 && j < no;
        if (!__any_sync(0xffffffffu, valid)) continue;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (valid) { rx = others[3 * j] - cx; ry = others[3 * j + 1] - cy; rz = others[3 * j + 2] - cz; }
#pragma unroll
        for (int k = 0; k < 15; k++) {
            const float w = valid ? influence_w(rx, ry, rz, s_kp[3 * k], s_kp[3 * k + 1], s_kp[3 * k + 2], inv_ext) : 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, w > 0.f);
            if (w > 0.f) {
                int2 e;
                e.x = (int)((unsigned)j | ((unsigned)k << K_SHIFT));
                e.y = __float_as_int(w);
                my_entries[run[k] + __popc(m & lt_mask)] = e;
            }
            run[k] += __popc(m);
        }
    }
}

// One warp per centre. The real neighbours of the row are compacted into shared memory (shadow entries dropped), then
// the warp sweeps the (kernel point, neighbour) pairs in kernel-point-major order, 32 pairs per step, one influence
// evaluation per lane: ballot-compacting the non-zero weights in that order yields the list already grouped by kernel
// point and ordered by table column, in a single pass.
__global__ void __launch_bounds__(INF_WARPS * 32) kp_influence_kernel(const float* __restrict__ centres, int nc,
                                                                     const float* __restrict__ others, int no, Table T,
                                                                     const float* __restrict__ kp, int K, float kp_sign,
                                                                     float inv_ext, unsigned short* __restrict__ koff,
                                                                     int2* __restrict__ entries) {
    __shared__ float s_kp[16 * 3];
    __shared__ float4 s_nb[INF_WARPS][INF_MAX_ROW];
    if (threadIdx.x < 48) s_kp[threadIdx.x] = (threadIdx.x < 3 * K) ? kp_sign * kp[threadIdx.x] : 1e30f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * INF_WARPS + warp;
    if (i >= nc) return;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float cx = centres[3 * (size_t)i], cy = centres[3 * (size_t)i + 1], cz = centres[3 * (size_t)i + 2];
    size_t row0, pos0;
    int cnt_row;
    if (T.rowptr) { row0 = (size_t)T.rowptr[i]; pos0 = row0; cnt_row = T.rowptr[i + 1] - T.rowptr[i]; }
    else { row0 = (size_t)i * T.H; pos0 = (size_t)i * T.stride; cnt_row = T.H; }
    int2* my_entries = entries + 15 * row0;
    unsigned short* koff_row = koff + (size_t)i * KOFF;
    if (cnt_row > INF_MAX_ROW) return;  // kp_influence_long_kernel takes these rows
    float4* nb = s_nb[warp];
    int hn = 0;
    for (int hb = 0; hb < cnt_row; hb += 32) {
        const int h = hb + lane;
        long long j = (h < cnt_row) ? table_get(T, pos0 + h) : -1;
        const bool valid = j >= 0 && j < no;
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid)
            nb[hn + __popc(m & lt_mask)] = make_float4(others[3 * j] - cx, others[3 * j + 1] - cy, others[3 * j + 2] - cz,
                                                       __int_as_float((int)j));
        hn += __popc(m);
    }
    __syncwarp();
    const int npairs = K * hn;
    const int my_first = min(lane * hn, npairs);  // first pair of kernel point `lane` (lanes 0..15)
    int myoff = 0, run = 0;
    // pair p = (kernel point p / hn, neighbour p % hn). hn <= 128 and p < 1920, so the quotient comes exactly from one
    // float multiply: (p + 0.5) / hn stays at least 0.5 / 128 away from an integer, far above the rounding error
    const float inv_hn = hn > 0 ? 1.f / (float)hn : 0.f;
    int base = 0;
    for (; base < npairs; base += 32) {
        const int p = base + lane;
        const bool valid = p < npairs;
        const int k = (int)(((float)p + 0.5f) * inv_hn);
        const int h = p - k * hn;
        float w = 0.f;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            v = nb[h];
            w = influence_w(v.x, v.y, v.z, s_kp[3 * k], s_kp[3 * k + 1], s_kp[3 * k + 2], inv_ext);
        }
        const unsigned m = __ballot_sync(0xffffffffu, w > 0.f);
        if (w > 0.f) {
            int2 e;
            e.x = (int)((unsigned)__float_as_int(v.w) | ((unsigned)k << K_SHIFT));
            e.y = __float_as_int(w);
            my_entries[run + __popc(m & lt_mask)] = e;
        }
        if (my_first >= base && my_first < base + 32) myoff = run + __popc(m & ((1u << (my_first - base)) - 1u));
        run += __popc(m);
    }
    if (my_first >= base) myoff = run;  // kernel points that start at or after the end of the sweep
    if (lane < 16) koff_row[lane] = (unsigned short)myoff;
}

// rows longer than INF_MAX_ROW (only possible for very dense tables); every other warp exits at once
__global__ void __launch_bounds__(INF_WARPS * 32) kp_influence_long_kernel(const float* __restrict__ centres, int nc,
                                                                          const float* __restrict__ others, int no,
                                                                          Table T, const float* __restrict__ kp, int K,
                                                                          float kp_sign, float inv_ext,
                                                                          unsigned short* __restrict__ koff,
                                                                          int2* __restrict__ entries) {
    if (T.rowptr && T.rowptr[nc + 1] <= INF_MAX_ROW) return;  // CSR tables record their longest row after the pointers
    __shared__ float s_kp[16 * 3];
    if (threadIdx.x < 48) s_kp[threadIdx.x] = (threadIdx.x < 3 * K) ? kp_sign * kp[threadIdx.x] : 1e30f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * INF_WARPS + (threadIdx.x >> 5);
    if (i >= nc) return;
    size_t row0, pos0;
    int cnt_row;
    if (T.rowptr) { row0 = (size_t)T.rowptr[i]; pos0 = row0; cnt_row = T.rowptr[i + 1] - T.rowptr[i]; }
    else { row0 = (size_t)i * T.H; pos0 = (size_t)i * T.stride; cnt_row = T.H; }
    if (cnt_row <= INF_MAX_ROW) return;
    influence_long_row(others, no, T, pos0, cnt_row, centres[3 * (size_t)i], centres[3 * (size_t)i + 1],
                       centres[3 * (size_t)i + 2], s_kp, inv_ext, koff + (size_t)i * KOFF, entries + 15 * row0, lane);
}

// ---------------------------------------------------------------------------------------------------- weight pack
// images[chunk][nblk] : NB rows (output channels) x CK reduction columns, K-major core-matrix layout:
//   element (n, col) at (n/8)*128 + (col/4)*(NB*16) + (n%8)*16 + (col%4)*4 bytes.  col <-> (k, c) = (col / cin_p, col % cin_p)
// value = tf32(W[k*sk + c*sc + n*sn]) inside the valid range, else 0.
__global__ void __launch_bounds__(256) kp_pack_w_kernel(const float* __restrict__ W, int K, int cin, int cin_p, int cout,
                                                       long long sk, long long sc, long long sn, int NB, int n_nblk,
                                                       int n_chunks, float* __restrict__ images) {
    const long long total = (long long)n_chunks * n_nblk * NB * CK;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        // t enumerates destination floats linearly (coalesced stores)
        const long long img = t / ((long long)NB * CK);
        const int r = (int)(t % ((long long)NB * CK));
        const int chunk = (int)(img / n_nblk), nblk = (int)(img % n_nblk);
        const int j = r / (NB * 4);          // 16-byte K chunk
        const int rem = r % (NB * 4);
        const int n8 = rem / 32, in8 = rem % 32;
        const int n = n8 * 8 + in8 / 4, e = in8 % 4;
        const int col = chunk * CK + j * 4 + e;
        const int k = col / cin_p, c = col % cin_p;
        const int ng = nblk * NB + n;
        float v = 0.f;
        if (k < K && c < cin && ng < cout) v = to_tf32(W[k * sk + c * sc + ng * sn]);
        images[t] = v;
    }
}

// zero-pad the channel dimension to a multiple of 4 (float4 gathers)
__global__ void __launch_bounds__(256) kp_pad_cols_kernel(const float* __restrict__ src, long long rows, int c, int c_p,
                                                         float* __restrict__ dst) {
    const long long total = rows * c_p;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / c_p;
        const int cc = (int)(t % c_p);
        dst[t] = cc < c ? src[r * c + cc] : 0.f;
    }
}

// ---------------------------------------------------------------------------------- A-tile assembly (shared by fwd, dW)
// Where the 16 bytes (4 consecutive reduction columns 4*lane..4*lane+3) of tile row p live in shared memory.
struct LayoutKMajor {  // forward: UMMA K-major, no swizzle (see A_SBO / A_LBO above)
    static __device__ __forceinline__ int off(int p, int lane) { return lane * A_LBO + (p >> 3) * A_SBO + (p & 7) * 16; }
};
// dW: the same tile consumed MN-major (M = reduction column, K = point). For 32-bit operands the only MN-major shared
// memory layout the tensor core accepts is SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom): rows of 32
// elements (128 B) along M, 4 consecutive K values = 4 consecutive rows (512 B atom), byte-address bits [5,7) XORed
// with bits [7,9). Element (m, k) at
//   (m/32)*MN_LBO + (k/4)*MN_SBO + (k%4)*128 + ((((m%32)/8) ^ (k%4))*32) + (m%8)*4
constexpr int MN_SBO = 512;                     // next group of 4 K values (points)
constexpr int MN_LBO = (TILE_M / 4) * MN_SBO;   // next group of 32 M values: 16 KiB
struct LayoutMNMajor {
    static __device__ __forceinline__ int off(int p, int lane) {
        return (lane >> 3) * MN_LBO + (p >> 2) * MN_SBO + (p & 3) * 128 + ((((lane & 7) >> 1) ^ (p & 3)) << 5) + (lane & 1) * 16;
    }
};

// Warp `warp` owns tile rows [warp*RPW, warp*RPW + RPW). For chunk `chunk` it zeroes them, walks the entry lists of
// its points restricted to the chunk's kernel points (flattened across the points, 32 entries per batch), gathers
// float4 feature rows with U loads in flight per lane and accumulates w * x into the rows, then rounds them to TF32.
template <int U, class LAY, int RPW>
__device__ __forceinline__ void assemble_rows(unsigned char* sA, int warp, int lane, int chunk, int cin_p, int K,
                                              const int* s_row0, const unsigned short* s_koff,
                                              const int2* __restrict__ entries, const float* __restrict__ x) {
    const int col0 = chunk * CK;
    const int kfirst = col0 / cin_p;
    const int klast = min(K - 1, (col0 + CK - 1) / cin_p);
    const int my_col = col0 + 4 * lane;
    const int k_l = my_col / cin_p, c_l = my_col % cin_p;
    const int p0 = warp * RPW;
#pragma unroll
    for (int r = 0; r < RPW; r++)
        *reinterpret_cast<float4*>(sA + LAY::off(p0 + r, lane)) = make_float4(0.f, 0.f, 0.f, 0.f);
    long long start = 0;
    int cnt = 0;
    if (lane < RPW && kfirst < K) {
        const unsigned short* ko = s_koff + (p0 + lane) * KOFF;
        start = 15LL * s_row0[p0 + lane] + ko[kfirst];
        cnt = (int)ko[klast + 1] - (int)ko[kfirst];
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < RPW; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int E = __shfl_sync(0xffffffffu, incl, RPW - 1);
    const int excl = incl - cnt;
    const float* xc = x + c_l;
    const int kc = klast - kfirst + 1;  // kernel points in this chunk (lanes with equal k_l form a group)
    for (int b0 = 0; b0 < E; b0 += 32) {
        const int e = b0 + lane;
        int pt = 0;
#pragma unroll
        for (int t = 0; t < RPW - 1; t++) pt += (__shfl_sync(0xffffffffu, incl, t) <= e) ? 1 : 0;
        const long long pstart = __shfl_sync(0xffffffffu, start, pt);
        const int pexcl = __shfl_sync(0xffffffffu, excl, pt);
        int2 rec = make_int2(0, 0);
        int ek = -1;
        if (e < E) {
            rec = entries[pstart + (e - pexcl)];
            ek = (int)((unsigned)rec.x >> K_SHIFT);
            rec.x = (int)(((unsigned)rec.x & J_MASK) | ((unsigned)pt << K_SHIFT));  // the kernel point is implied by the
        }                                                                          // consuming lane group: carry the row
        // Each lane group (lanes sharing a kernel point) walks ITS entries of the batch; groups advance together, so
        // one step serves up to kc entries.
        unsigned mymask = 0;
        for (int g = 0; g < kc; g++) {
            const unsigned m = __ballot_sync(0xffffffffu, ek == kfirst + g);
            if (k_l == kfirst + g) mymask = m;
        }
        int steps = __popc(mymask);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        for (int g = 0; g < steps; g += U) {
            float4 xv[U];
            float wv[U];
            int ov[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool have = mymask != 0u;
                const int ee = have ? (__ffs(mymask) - 1) : 0;
                mymask &= mymask - 1u;
                const unsigned jp = (unsigned)__shfl_sync(0xffffffffu, rec.x, ee);
                const float w = __int_as_float(__shfl_sync(0xffffffffu, rec.y, ee));
                ov[u] = LAY::off(p0 + (int)(jp >> K_SHIFT), lane);  // (row, lane) are coupled by the swizzle
                wv[u] = have ? w : 0.f;
                xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) xv[u] = __ldg(reinterpret_cast<const float4*>(xc + (size_t)(jp & J_MASK) * cin_p));
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (wv[u] != 0.f) {
                    float4* a = reinterpret_cast<float4*>(sA + ov[u]);
                    float4 v = *a;
                    v.x = fmaf(wv[u], xv[u].x, v.x); v.y = fmaf(wv[u], xv[u].y, v.y);
                    v.z = fmaf(wv[u], xv[u].z, v.z); v.w = fmaf(wv[u], xv[u].w, v.w);
                    *a = v;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        float4* a = reinterpret_cast<float4*>(sA + LAY::off(p0 + r, lane));
        float4 v = *a;
        v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
        *a = v;
    }
}

// Dense variant of the A tile (linear layers next to KPConv: the A operand is a plain row-major matrix): warp `warp`
// copies its RPW rows of columns [chunk*CK, chunk*CK + CK) from a[n, ld], optionally scaled by the LeakyReLU derivative
// taken from `mask` (same shape: factor 1 where mask > 0, `slope` elsewhere), rounded to TF32. All RPW loads of a lane are
// independent and issued together.
template <class LAY, int RPW>
__device__ __forceinline__ void dense_rows(unsigned char* sA, int warp, int lane, int chunk, int tile_base, int n,
                                           const float* __restrict__ a, int ld, const float* __restrict__ mask,
                                           float slope) {
    const int col = chunk * CK + 4 * lane;
    const int p0 = warp * RPW;
    float4 v[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        const int i = tile_base + p0 + r;
        v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n && col < ld) v[r] = __ldg(reinterpret_cast<const float4*>(a + (size_t)i * ld + col));
    }
    if (mask) {
#pragma unroll
        for (int r = 0; r < RPW; r++) {
            const int i = tile_base + p0 + r;
            if (i < n && col < ld) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(mask + (size_t)i * ld + col));
                v[r].x *= y.x > 0.f ? 1.f : slope; v[r].y *= y.y > 0.f ? 1.f : slope;
                v[r].z *= y.z > 0.f ? 1.f : slope; v[r].w *= y.w > 0.f ? 1.f : slope;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        v[r].x = to_tf32(v[r].x); v[r].y = to_tf32(v[r].y); v[r].z = to_tf32(v[r].z); v[r].w = to_tf32(v[r].w);
        *reinterpret_cast<float4*>(sA + LAY::off(p0 + r, lane)) = v[r];
    }
}

// Second formulation of the same assembly (the default, KP_ASSEMBLE_V=2). Every lane owns 4 reduction columns of the
// tile, i.e. one kernel point k_l and 4 channels, and the entries it needs from row r are exactly the k_l group of that
// row's list: entries[15*row0(r) + koff[r][k_l] .. koff[r][k_l+1]). So each lane walks its OWN short entry stream and
// accumulates in registers: no zero pass over the tile, no read-modify-write of shared memory, no second rounding pass
// and none of the ballot / ffs / shuffle bookkeeping that distributes a flat entry list over lane groups (that version
// executes ~2x the instructions). Rows are processed R = 8 at a time so that 8 independent gathers are in flight per
// lane, and the entry record of step t+1 is requested while the feature rows of step t are in flight. Lanes that share
// a kernel point read the same record (one broadcast transaction).
template <int U, class LAY, int RPW>
__device__ __forceinline__ void assemble_rows_v2(unsigned char* sA, int warp, int lane, int chunk, int cin_p, int K,
                                                 const int* s_row0, const unsigned short* s_koff,
                                                 const int2* __restrict__ entries, const float* __restrict__ x) {
    constexpr int R = 8;
    static_assert(RPW % R == 0, "rows per warp must be a multiple of the row group");
    const int my_col = chunk * CK + 4 * lane;
    const int k_l = my_col / cin_p, c_l = my_col - k_l * cin_p;
    const bool lane_ok = k_l < K;
    const float* xc = x + c_l;
    const int p0 = warp * RPW;
#pragma unroll 1
    for (int r0 = 0; r0 < RPW; r0 += R) {
        const int2* ep[R];
        int cnt[R];
        float4 acc[R];
        int2 nxt[R];
        int m = 0;
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int row = p0 + r0 + u;
            int b = 0, e = 0;
            if (lane_ok) {
                const unsigned short* ko = s_koff + row * KOFF + k_l;
                b = ko[0];
                e = ko[1];
            }
            ep[u] = entries + 15LL * s_row0[row] + b;
            cnt[u] = e - b;
            m = max(m, cnt[u]);
            acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            nxt[u] = make_int2(0, 0);
            if (cnt[u] > 0) nxt[u] = __ldg(ep[u]);
        }
        for (int t = 0; t < m; t++) {
            float4 xv[R];
            float wv[R];
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int2 rec = nxt[u];
                xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                wv[u] = 0.f;
                if (t < cnt[u]) {
                    xv[u] = __ldg(reinterpret_cast<const float4*>(xc + (size_t)((unsigned)rec.x & J_MASK) * cin_p));
                    wv[u] = __int_as_float(rec.y);
                }
                if (t + 1 < cnt[u]) nxt[u] = __ldg(ep[u] + t + 1);
            }
#pragma unroll
            for (int u = 0; u < R; u++) {
                acc[u].x = fmaf(wv[u], xv[u].x, acc[u].x); acc[u].y = fmaf(wv[u], xv[u].y, acc[u].y);
                acc[u].z = fmaf(wv[u], xv[u].z, acc[u].z); acc[u].w = fmaf(wv[u], xv[u].w, acc[u].w);
            }
        }
#pragma unroll
        for (int u = 0; u < R; u++) {
            float4 v = acc[u];
            v.x = to_tf32(v.x); v.y = to_tf32(v.y); v.z = to_tf32(v.z); v.w = to_tf32(v.w);
            *reinterpret_cast<float4*>(sA + LAY::off(p0 + r0 + u, lane)) = v;
        }
    }
}

#ifndef KP_ASSEMBLE_V
#define KP_ASSEMBLE_V 2
#endif
#if KP_ASSEMBLE_V == 1
#define KP_ASSEMBLE assemble_rows
#else
#define KP_ASSEMBLE assemble_rows_v2
#endif

// stage the entry-list headers of one tile: row0 (first table column of the centre) and koff (32 bytes per centre,
// moved as two 16-byte words)
template <int FWD_THREADS>
__device__ __forceinline__ void stage_headers(int tile_base, int n, int H, const int* __restrict__ rowptr,
                                              const unsigned short* __restrict__ koff, int* s_row0,
                                              unsigned short* s_koff) {
    for (int t = threadIdx.x; t < TILE_M; t += FWD_THREADS) {
        const int i = tile_base + t;
        s_row0[t] = (i < n) ? (rowptr ? rowptr[i] : i * H) : 0;
    }
    const uint4* src = reinterpret_cast<const uint4*>(koff);
    uint4* dst = reinterpret_cast<uint4*>(s_koff);
    for (int t = threadIdx.x; t < TILE_M * 2; t += FWD_THREADS) {
        const int i = tile_base + (t >> 1);
        dst[t] = (i < n) ? __ldg(src + (size_t)i * 2 + (t & 1)) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// --------------------------------------------------------------------------------------------------------- forward
struct FwdParams {
    int nq;              // centre points (rows of out)
    const float* x;      // [n_other, cin_p] features gathered through the entry lists
    int cin_p, K;
    const int* rowptr;   // CSR row pointers of the centres' table, or null for a padded table of width H
    int H;
    const unsigned short* koff;
    const int2* entries;
    const float* images; // packed weights [n_chunks][n_nblk][NB*CK]
    int NB, n_nblk, n_chunks;
    int ksplit;          // CTAs along the reduction (blockIdx.y); > 1 => partial sums are added atomically into out
    float* out;          // [nq, cout] with row stride ldo (pre-zeroed when ksplit > 1)
    int cout, ldo;
    uint32_t tmem_cols;
    // dense mode (template DENSE): A = x[nq, cin_p] itself, optionally times the LeakyReLU derivative read from mask
    const float* mask;
    float slope_in;
    // epilogue (ksplit == 1 only): out = leaky(acc + bias, slope_out); bias may be null, slope_out = 1 disables
    const float* bias;
    float slope_out;
};

template <int NW, bool DENSE>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) kp_fwd_kernel(FwdParams P) {
    constexpr int FWD_THREADS = NW * 32, RPW = TILE_M / NW, NWARPS = NW;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + A_BYTES;
    const int b_bytes = P.NB * CK * 4;
    unsigned short* s_koff = reinterpret_cast<unsigned short*>(sB + b_bytes);
    int* s_row0 = reinterpret_cast<int*>(s_koff + TILE_M * KOFF);
    uint64_t* bar_b = reinterpret_cast<uint64_t*>(s_row0 + TILE_M);
    uint64_t* bar_mma = bar_b + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_base = blockIdx.x * TILE_M;
    const int cps = (P.n_chunks + P.ksplit - 1) / P.ksplit;
    const int c0 = blockIdx.y * cps, c1 = min(P.n_chunks, c0 + cps);
    if (c0 >= c1) return;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(s_tmem, P.tmem_cols);
    if (!DENSE) stage_headers<FWD_THREADS>(tile_base, P.nq, P.H, P.rowptr, P.koff, s_row0, s_koff);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = make_idesc(TILE_M, P.NB, 0, 0);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    const uint32_t b_lbo = (uint32_t)P.NB * 16u;

    int step = 0;  // one MMA group (chunk, nblk) per step; bar_b / bar_mma complete once per step
    for (int chunk = c0; chunk < c1; chunk++) {
        for (int nblk = 0; nblk < P.n_nblk; nblk++, step++) {
            if (step > 0) mbar_wait(bar_mma, (uint32_t)((step - 1) & 1));  // previous MMAs done: A and B reusable
            if (tid == 0) {
                mbar_expect_tx(bar_b, (uint32_t)b_bytes);
                bulk_g2s(sB, P.images + ((size_t)chunk * P.n_nblk + nblk) * (size_t)P.NB * CK, (uint32_t)b_bytes, bar_b);
            }
            if (nblk == 0) {
                if (DENSE) dense_rows<LayoutKMajor, RPW>(sA, warp, lane, chunk, tile_base, P.nq, P.x, P.cin_p, P.mask, P.slope_in);
                else KP_ASSEMBLE<8, LayoutKMajor, RPW>(sA, warp, lane, chunk, P.cin_p, P.K, s_row0, s_koff, P.entries, P.x);
                fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
            }
            __syncthreads();
            if (tid == 0) {
                mbar_wait(bar_b, (uint32_t)(step & 1));
                tc_fence_after();
#pragma unroll 1
                for (int kk = 0; kk < CK / 8; kk++) {  // K = 8 per tf32 MMA = two 16-byte K chunks
                    const uint64_t ad = make_desc(a_addr + kk * 2 * A_LBO, A_LBO, A_SBO);
                    const uint64_t bd = make_desc(b_addr + kk * 2 * b_lbo, b_lbo, B_SBO);
                    umma_tf32(tmem + (uint32_t)(nblk * P.NB), ad, bd, idesc, (chunk > c0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(bar_mma);
            }
        }
    }
    mbar_wait(bar_mma, (uint32_t)((step - 1) & 1));
    tc_fence_after();

    // epilogue: warp w reads TMEM lanes 32*(w%4).., column blocks of 16 dealt round-robin over the 4 warps of a quadrant
    const int row = tile_base + 32 * (warp & 3) + lane;
    const int n_cb = (P.n_nblk * P.NB) / 16;
    const bool vec = (P.cout & 3) == 0 && (P.ldo & 3) == 0;
    const bool post = P.ksplit == 1 && (P.bias != nullptr || P.slope_out != 1.f);
    for (int cb = warp >> 2; cb < n_cb; cb += NWARPS / 4) {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------- dW
// B tile = dOut[tile points, NB outputs], also MN-major SWIZZLE_128B_BASE32B: element (n, p) by the formula above
// with m = n. Size = ceil(NB/32) * MN_LBO.
__host__ __device__ constexpr int dw_b_bytes(int NB) { return ((NB + 31) / 32) * MN_LBO; }

struct DwParams {
    int nq;
    const float* x;      // [ns, cin_p]
    int cin, cin_p, K, H;
    const unsigned short* koff;
    const int2* entries;
    const float* dout;   // [nq, cout]
    int cout, NB;        // NB = output columns handled per CTA (<= 256), slice index = blockIdx.z
    int n_tiles, n_splits;
    float* dw;           // [K, cin, cout], pre-zeroed
    uint32_t tmem_cols;
    // dense mode: A = x[nq, cin_p] itself (times the LeakyReLU derivative read from mask), see dense_rows
    const float* mask;
    float slope_in;
};

template <int NW, bool DENSE>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) kp_dw_kernel(DwParams P) {
    constexpr int FWD_THREADS = NW * 32, RPW = TILE_M / NW, NWARPS = NW;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                       // 4 * MN_LBO = 64 KiB, 1024-aligned (swizzle uses address bits)
    unsigned char* sB = smem + 4 * MN_LBO;
    const int b_bytes = dw_b_bytes(P.NB);
    unsigned short* s_koff = reinterpret_cast<unsigned short*>(sB + b_bytes);
    int* s_row0 = reinterpret_cast<int*>(s_koff + TILE_M * KOFF);
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(s_row0 + TILE_M);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, split = blockIdx.y, n0 = blockIdx.z * P.NB;
    if (split >= P.n_tiles) return;  // uniform: nothing to do for this CTA

    if (tid == 0) {
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(s_tmem, P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t idesc = make_idesc(TILE_M, P.NB, 1, 1);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);

    int step = 0;
    for (int tile = split; tile < P.n_tiles; tile += P.n_splits, step++) {
        const int tile_base = tile * TILE_M;
        if (step > 0) mbar_wait(bar_mma, (uint32_t)((step - 1) & 1));
        if (!DENSE) stage_headers<FWD_THREADS>(tile_base, P.nq, P.H, nullptr, P.koff, s_row0, s_koff);
        // dOut tile -> B (TF32): the warp's RPW rows x NB/4 float4 items are dealt round-robin to the lanes, eight loads
        // in flight per lane (a row-at-a-time loop costs RPW dependent global-memory round trips per tile)
        {
            const int nv = P.NB >> 2;
            const int items = RPW * nv;
            const bool vec4 = (P.cout & 3) == 0;
            for (int base = 0; base < items; base += 32 * 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int t = base + u * 32 + lane;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t < items) {
                        const int r = t / nv, n4 = t - r * nv;
                        const int i = tile_base + warp * RPW + r;
                        const int n = n0 + n4 * 4;
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
                for (int u = 0; u < 8; u++) {
                    const int t = base + u * 32 + lane;
                    if (t < items) {
                        const int r = t / nv, n4 = t - r * nv;
                        float4 w = v[u];
                        w.x = to_tf32(w.x); w.y = to_tf32(w.y); w.z = to_tf32(w.z); w.w = to_tf32(w.w);
                        *reinterpret_cast<float4*>(sB + LayoutMNMajor::off(warp * RPW + r, n4)) = w;
                    }
                }
            }
        }
        __syncthreads();  // headers ready
        if (DENSE) dense_rows<LayoutMNMajor, RPW>(sA, warp, lane, chunk, tile_base, P.nq, P.x, P.cin_p, P.mask, P.slope_in);
        else KP_ASSEMBLE<8, LayoutMNMajor, RPW>(sA, warp, lane, chunk, P.cin_p, P.K, s_row0, s_koff, P.entries, P.x);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll 1
            for (int kk = 0; kk < TILE_M / 8; kk++) {  // K = 8 points per MMA
                // MN-major swizzled descriptors: "leading" offset = next 32-element group along M/N, "stride" offset =
                // next group of 4 K values; 8 points per MMA = two K groups = 1024 B
                const uint64_t ad = make_desc(a_addr + kk * 2 * MN_SBO, MN_LBO, MN_SBO, LAYOUT_SW128_BASE32B);
                const uint64_t bd = make_desc(b_addr + kk * 2 * MN_SBO, MN_LBO, MN_SBO, LAYOUT_SW128_BASE32B);
                umma_tf32(tmem, ad, bd, idesc, (step > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
    }
    mbar_wait(bar_mma, (uint32_t)((step - 1) & 1));
    tc_fence_after();

    // epilogue: TMEM lane r = reduction column chunk*CK + r = (k, c); add into dW[k, c, n0 + col]
    const int r = 32 * (warp & 3) + lane;
    const int col = chunk * CK + r;
    const int k = col / P.cin_p, c = col % P.cin_p;
    const bool row_ok = k < P.K && c < P.cin;
    const bool vec = (P.cout & 3) == 0;
    for (int cb = warp >> 2; cb < P.NB / 16; cb += NWARPS / 4) {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, P.tmem_cols);
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
struct Lists {
    unsigned short* koff;
    int2* entries;
};

static int build_lists(Scratch& S, const float* centres, int nc, const float* others, int no, const Table& T,
                       long long n_pairs, int max_row, const float* kp, int K, float kp_sign, float extent, Lists* L,
                       cudaStream_t stream, void* ext_koff = nullptr, void* ext_entries = nullptr) {
    if (n_pairs * 15 >= (1LL << 31)) return fail(KP_ERR_UNSUPPORTED, "kpconv: neighbour table too large (Nq*H*15 >= 2^31)");
    L->koff = ext_koff ? (unsigned short*)ext_koff : S.alloc<unsigned short>((size_t)nc * KOFF);
    L->entries = ext_entries ? (int2*)ext_entries : S.alloc<int2>((size_t)(n_pairs > 0 ? n_pairs : 1) * 15);
    if (S.status != KP_OK) return S.status;
    ProfileScope ps("kp_influence", stream);
    kp_influence_kernel<<<ceil_div(nc, INF_WARPS), INF_WARPS * 32, 0, stream>>>(centres, nc, others, no, T, kp, K, kp_sign,
                                                                             1.f / extent, L->koff, L->entries);
    if (max_row == 0 || max_row > INF_MAX_ROW) {  // CSR rows of unknown length, or a padded table wider than the staging
        KP_CHECK_LAUNCH();
        kp_influence_long_kernel<<<ceil_div(nc, INF_WARPS), INF_WARPS * 32, 0, stream>>>(centres, nc, others, no, T, kp, K,
                                                                                      kp_sign, 1.f / extent, L->koff,
                                                                                      L->entries);
    }
    KP_CHECK_LAUNCH();
    return KP_OK;
}

static int pad4(int c) { return (c + 3) & ~3; }

// How many CTAs share the reduction axis of one 128-point tile. A CTA costs (its chunks + ~1 chunk of fixed work:
// TMEM allocation, headers, epilogue), CTAs run in waves of `slots` (2 per SM for the 8-warp kernels, 1 for the 16-warp
// ones), and a split pays a zero-fill plus an atomic epilogue. The previous rule (fill 296 slots, never look at the
// wave count) ran the 146-tile layer as 438 CTAs = 1.5 waves of 3 chunks where 292 CTAs = 1 wave of 4 chunks is shorter,
// and split the 252-tile layer in two for nothing. WEASAL_KSPLIT_MODEL=0 restores it (A/B runs).
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


// opt-in dynamic shared memory: raise a kernel's limit only when a launch needs more than it already has
template <typename KernelT>
static cudaError_t set_smem(KernelT kernel, size_t bytes) {
    static thread_local std::map<std::pair<const void*, int>, size_t> have;  // (kernel, device) -> current limit
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& h = have[{(const void*)kernel, dev}];
    if (bytes <= h) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) h = bytes;
    return e;
}

// out[nc, cout] = sum over entry lists of w * x[j, :] contracted with W (strides sk, sc, sn over (k, c_in, n_out))
static int run_forward(const char* tag, Scratch& S, int nc, const int* rowptr, int H, const float* x, int n_x_rows, int cin,
                       const Lists& L, const float* W, long long sk, long long sc, long long sn, int cout, int K,
                       float* out, cudaStream_t stream) {
    const int cin_p = pad4(cin);
    const float* xg = x;
    if (cin_p != cin) {
        float* xp = S.alloc<float>((size_t)n_x_rows * cin_p);
        if (S.status != KP_OK) return S.status;
        const long long tot = (long long)n_x_rows * cin_p;
        kp_pad_cols_kernel<<<ceil_div(tot, 256) < 2048 ? ceil_div(tot, 256) : 2048, 256, 0, stream>>>(x, n_x_rows, cin, cin_p, xp);
        KP_CHECK_LAUNCH();
        xg = xp;
    }
    const int cout_p = (cout + 15) & ~15;
    const int NB = cout_p < 256 ? cout_p : 256;
    const int n_nblk = ceil_div(cout_p, NB);
    if (n_nblk * NB > 512) return fail(KP_ERR_UNSUPPORTED, "kpconv: out_channels > 512");
    const int n_chunks = ceil_div((long long)K * cin_p, CK);
    float* images = S.alloc<float>((size_t)n_chunks * n_nblk * NB * CK);
    if (S.status != KP_OK) return S.status;
    {
        const long long total = (long long)n_chunks * n_nblk * NB * CK;
        const int grid = ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096;
        ProfileScope ps("kp_pack_w", stream);
        kp_pack_w_kernel<<<grid, 256, 0, stream>>>(W, K, cin, cin_p, cout, sk, sc, sn, NB, n_nblk, n_chunks, images);
        KP_CHECK_LAUNCH();
    }
    FwdParams P;
    P.nq = nc; P.x = xg; P.cin_p = cin_p; P.K = K;
    P.rowptr = rowptr; P.H = H;
    P.koff = L.koff; P.entries = L.entries;
    P.images = images; P.NB = NB; P.n_nblk = n_nblk; P.n_chunks = n_chunks;
    P.out = out; P.cout = cout; P.ldo = cout;
    P.mask = nullptr; P.slope_in = 1.f; P.bias = nullptr; P.slope_out = 1.f;
    // split the reduction across CTAs when the tiles alone cannot fill the 148 SMs (two CTAs each)
    const int n_tiles = ceil_div(nc, TILE_M);
    // Split partial sums meet in `out` through float atomics, so their order (the last bits of the result) varies from
    // run to run; WEASAL_KPCONV_DETERMINISTIC=1 keeps one CTA per tile.
    const size_t smem = (size_t)A_BYTES + (size_t)NB * CK * 4 + TILE_M * KOFF * 2 + TILE_M * 4 + 64;
    const int ksplit = pick_ksplit(n_tiles, n_chunks, smem <= (size_t)SMEM_TWO_CTAS ? 2 * 148 : 148);
    P.ksplit = ksplit;
    if (ksplit > 1) KP_CUDA(cudaMemsetAsync(out, 0, (size_t)nc * cout * sizeof(float), stream));
    uint32_t cols = 32;
    while ((int)cols < n_nblk * NB) cols <<= 1;
    P.tmem_cols = cols;
    if (smem <= (size_t)SMEM_TWO_CTAS) {
        KP_CUDA(set_smem(kp_fwd_kernel<8, false>, smem));
        ProfileScope ps(tag, stream);
        kp_fwd_kernel<8, false><<<dim3(n_tiles, ksplit), 256, smem, stream>>>(P);
    } else {
        KP_CUDA(set_smem(kp_fwd_kernel<16, false>, smem));
        ProfileScope ps(tag, stream);
        kp_fwd_kernel<16, false><<<dim3(n_tiles, ksplit), 512, smem, stream>>>(P);
    }
    KP_CHECK_LAUNCH();
    return KP_OK;
}

static int check_args(int nq, int ns, int H, int idx_stride, int cin, int cout, int K, float extent) {
    if (nq < 0 || ns < 0 || H < 0 || idx_stride < H || cin <= 0 || cout <= 0) return fail(KP_ERR_ARG, "kpconv: bad sizes");
    if (K <= 0 || K > 15) return fail(KP_ERR_UNSUPPORTED, "kpconv: kernel_size must be 1..15");
    if (!(extent > 0.f)) return fail(KP_ERR_ARG, "kpconv: KP_extent must be positive");
    if (H > 4096) return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 4096 neighbour columns");
    if ((long long)ns >= (1LL << K_SHIFT) || (long long)nq >= (1LL << K_SHIFT))
        return fail(KP_ERR_UNSUPPORTED, "kpconv: more than 2^27 points in one call");
    return KP_OK;
}

// lists_koff / lists_entries (optional, caller-owned device buffers of kpconv_lists_bytes): the forward pass leaves
// the influence entry lists there so that backward can reuse them instead of rebuilding them.
void kpconv_lists_bytes(int nq, int H, long long* koff_bytes, long long* entries_bytes) {
    *koff_bytes = (long long)(nq > 0 ? nq : 1) * KOFF * 2;
    *entries_bytes = (long long)(nq > 0 ? nq : 1) * (H > 0 ? H : 1) * 15 * 8;
}

int kpconv_forward_device(const float* q, int nq, const float* s, int ns, const void* idx, int idx_is_i64, int H,
                          int idx_stride, const float* x, int cin, const float* w, int cout, const float* kp, int K,
                          float extent, float* out, void* lists_koff, void* lists_entries, cudaStream_t stream) {
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
    rc = build_lists(S, q, nq, s, ns, T, (long long)nq * H, H, kp, K, 1.f, extent, &L, stream, lists_koff, lists_entries);
    if (rc != KP_OK) return rc;
    return run_forward("kp_fwd", S, nq, nullptr, H, x, ns, cin, L, w, (long long)cin * cout, cout, 1, cout, K, out, stream);
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
                           float extent, const float* dout, float* dx, float* dw, const void* lists_koff,
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

    // ---- dW: entry lists centred on the queries (same as forward)
    {
        Table T;
        T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
        Lists L;
        if (lists_koff && lists_entries) {  // left behind by the forward pass
            L.koff = (unsigned short*)lists_koff;
            L.entries = (int2*)lists_entries;
        } else {
            rc = build_lists(S, q, nq, s, ns, T, n_pairs, H, kp, K, 1.f, extent, &L, stream);
            if (rc != KP_OK) return rc;
        }
        const int cin_p = pad4(cin);
        const float* xg = x;
        if (cin_p != cin) {
            float* xp = S.alloc<float>((size_t)ns * cin_p);
            if (S.status != KP_OK) return S.status;
            const int grid = ceil_div((long long)ns * cin_p, 256) < 2048 ? ceil_div((long long)ns * cin_p, 256) : 2048;
            kp_pad_cols_kernel<<<grid, 256, 0, stream>>>(x, ns, cin, cin_p, xp);
            KP_CHECK_LAUNCH();
            xg = xp;
        }
        const int cout_p = (cout + 15) & ~15;
        DwParams P;
        P.nq = nq; P.x = xg; P.cin = cin; P.cin_p = cin_p; P.K = K; P.H = H;
        P.koff = L.koff; P.entries = L.entries;
        P.dout = dout; P.cout = cout;
        P.NB = cout_p < 256 ? cout_p : 256;
        const int n_slices = ceil_div(cout_p, P.NB);
        const int n_chunks = ceil_div((long long)K * cin_p, CK);
        P.n_tiles = ceil_div(nq, TILE_M);
        const size_t smem = (size_t)4 * MN_LBO + (size_t)dw_b_bytes(P.NB) + TILE_M * KOFF * 2 + TILE_M * 4 + 64;
        int splits = (smem <= (size_t)SMEM_TWO_CTAS ? 2 * 148 : 148) / (n_chunks * n_slices);  // one wave of CTAs
        if (splits < 1) splits = 1;
        if (splits > P.n_tiles) splits = P.n_tiles;
        P.n_splits = splits;
        P.dw = dw;
        P.mask = nullptr; P.slope_in = 1.f;
        uint32_t cols = 32;
        while ((int)cols < P.NB) cols <<= 1;
        P.tmem_cols = cols;
        if (smem <= (size_t)SMEM_TWO_CTAS) {
            KP_CUDA(set_smem(kp_dw_kernel<8, false>, smem));
            ProfileScope ps("kp_dw", stream);
            kp_dw_kernel<8, false><<<dim3(n_chunks, splits, n_slices), 256, smem, stream>>>(P);
        } else {
            KP_CUDA(set_smem(kp_dw_kernel<16, false>, smem));
            ProfileScope ps("kp_dw", stream);
            kp_dw_kernel<16, false><<<dim3(n_chunks, splits, n_slices), 512, smem, stream>>>(P);
        }
        KP_CHECK_LAUNCH();
    }

    // ---- dX: the forward kernel on the transposed table, with W^T and -kp
    if (table_symmetric) {
        // queries == supports and no row was cropped: j is in row i exactly when i is in row j (the f32 distance is
        // exactly symmetric), so the table is its own transpose and no CSR copy is needed
        Table T;
        T.idx = idx; T.rowptr = nullptr; T.H = H; T.stride = idx_stride; T.is_i64 = idx_is_i64;
        Lists L;
        rc = build_lists(S, s, ns, q, nq, T, n_pairs, H, kp, K, -1.f, extent, &L, stream);
        if (rc != KP_OK) return rc;
        rc = run_forward("kp_fwd_dx", S, ns, nullptr, H, dout, nq, cout, L, w, (long long)cin * cout, 1, cout, cin, K, dx, stream);
        if (rc != KP_OK) return rc;
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
        Table T;
        T.idx = col_sorted; T.rowptr = rowptr; T.H = 0; T.stride = 0; T.is_i64 = 0;
        Lists L;
        rc = build_lists(S, s, ns, q, nq, T, n_pairs, 0, kp, K, -1.f, extent, &L, stream);
        if (rc != KP_OK) return rc;
        // W'[k][c' = o][n' = c] = W[k][c][o]
        rc = run_forward("kp_fwd_dx", S, ns, rowptr, 0, dout, nq, cout, L, w, (long long)cin * cout, 1, cout, cin, K, dx, stream);
        if (rc != KP_OK) return rc;
    }
    return KP_OK;
}

// host-side planning, exported for tests (no device work)
int plan_ksplit(int n_tiles, int n_chunks, int slots) { return pick_ksplit(n_tiles, n_chunks, slots); }

// ------------------------------------------------------------------------------------------ dense linear layers
// The unary blocks around every KPConv (models/blocks.py:467-507 UnaryBlock: Linear without bias -> BatchNorm, which
// is the identity on these 2-D features or a bias when use_bn is off, blocks.py:453-465 -> LeakyReLU(0.1)) run on the same
// tcgen05 pipeline with a dense A tile: one kernel for y = leaky(x W^T + b), one for dx = (dy * leaky'(y)) W, one for
// dW = (dy * leaky'(y))^T x. The library GEMMs they replace pick 9-18 CTAs for the weight gradients of the shallow
// layers (a 40k-row reduction into a 16x32 matrix) and need separate activation / bias kernels.
__global__ void __launch_bounds__(256) kp_bias_act_kernel(float* __restrict__ y, long long n, int cout, int ld,
                                                         const float* __restrict__ bias, float slope) {
    const long long total = n * cout;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cout;
        const int c = (int)(t % cout);
        float a = y[r * ld + c] + (bias ? bias[c] : 0.f);
        y[r * ld + c] = a > 0.f ? a : a * slope;
    }
}

// out[n, cout] (row stride ldo) = leaky( (a ⊙ leaky'(mask))[n, acols] · B + bias ), B(c, o) = W[c*sc + o*sn]
static int run_dense(const char* tag, Scratch& S, int n, const float* a, int acols, int acols_valid, const float* mask,
                     float slope_in, const float* W, long long sc, long long sn, int cout, const float* bias,
                     float slope_out, float* out, int ldo, cudaStream_t stream) {
    const int n_tiles = ceil_div(n, TILE_M);
    const int n_chunks = ceil_div(acols, CK);
    for (int o0 = 0; o0 < cout; o0 += 512) {  // TMEM holds 512 accumulator columns
        const int co = cout - o0 < 512 ? cout - o0 : 512;
        const int cout_p = (co + 15) & ~15;
        const int NB = cout_p < 256 ? cout_p : 256;
        const int n_nblk = ceil_div(cout_p, NB);
        float* images = S.alloc<float>((size_t)n_chunks * n_nblk * NB * CK);
        if (S.status != KP_OK) return S.status;
        {
            const long long total = (long long)n_chunks * n_nblk * NB * CK;
            const int grid = ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096;
            ProfileScope ps("lin_pack_w", stream);
            kp_pack_w_kernel<<<grid, 256, 0, stream>>>(W + (long long)o0 * sn, 1, acols_valid, acols, co, 0, sc, sn, NB, n_nblk,
                                                       n_chunks, images);
            KP_CHECK_LAUNCH();
        }
        FwdParams P;
        P.nq = n; P.x = a; P.cin_p = acols; P.K = 1;
        P.rowptr = nullptr; P.H = 0; P.koff = nullptr; P.entries = nullptr;
        P.images = images; P.NB = NB; P.n_nblk = n_nblk; P.n_chunks = n_chunks;
        P.out = out + o0; P.cout = co; P.ldo = ldo;
        P.mask = mask; P.slope_in = slope_in;
        P.bias = bias ? bias + o0 : nullptr; P.slope_out = slope_out;
        const size_t smem = (size_t)A_BYTES + (size_t)NB * CK * 4 + TILE_M * KOFF * 2 + TILE_M * 4 + 64;
        const int ksplit = pick_ksplit(n_tiles, n_chunks, smem <= (size_t)SMEM_TWO_CTAS ? 2 * 148 : 148);
        P.ksplit = ksplit;
        if (ksplit > 1) KP_CUDA(cudaMemset2DAsync(out + o0, (size_t)ldo * 4, 0, (size_t)co * 4, n, stream));
        uint32_t cols = 32;
        while ((int)cols < n_nblk * NB) cols <<= 1;
        P.tmem_cols = cols;
        if (smem <= (size_t)SMEM_TWO_CTAS) {
            KP_CUDA(set_smem(kp_fwd_kernel<8, true>, smem));
            ProfileScope ps(tag, stream);
            kp_fwd_kernel<8, true><<<dim3(n_tiles, ksplit), 256, smem, stream>>>(P);
        } else {
            KP_CUDA(set_smem(kp_fwd_kernel<16, true>, smem));
            ProfileScope ps(tag, stream);
            kp_fwd_kernel<16, true><<<dim3(n_tiles, ksplit), 512, smem, stream>>>(P);
        }
        KP_CHECK_LAUNCH();
        if (ksplit > 1 && (bias || slope_out != 1.f)) {  // the split partial sums met through atomics: finish separately
            const long long total = (long long)n * co;
            const int grid = ceil_div(total, 256) < 2368 ? ceil_div(total, 256) : 2368;
            kp_bias_act_kernel<<<grid, 256, 0, stream>>>(out + o0, n, co, ldo, P.bias, slope_out);
            KP_CHECK_LAUNCH();
        }
    }
    return KP_OK;
}

// rows padded to a multiple of 4 columns when needed (float4 loads of the dense A tile)
static int padded_cols(Scratch& S, const float* src, int n, int c, const float** dst, int* c_p, cudaStream_t stream) {
    *c_p = pad4(c);
    *dst = src;
    if (*c_p == c) return KP_OK;
    float* p = S.alloc<float>((size_t)n * *c_p);
    if (S.status != KP_OK) return S.status;
    const long long tot = (long long)n * *c_p;
    kp_pad_cols_kernel<<<ceil_div(tot, 256) < 2048 ? ceil_div(tot, 256) : 2048, 256, 0, stream>>>(src, n, c, *c_p, p);
    KP_CHECK_LAUNCH();
    *dst = p;
    return KP_OK;
}

static int check_linear(int n, int cin, int cout, float slope) {
    if (n < 0 || cin <= 0 || cout <= 0) return fail(KP_ERR_ARG, "linear: bad sizes");
    if (!(slope == slope)) return fail(KP_ERR_ARG, "linear: bad slope");
    return KP_OK;
}

// y[n, cout] = leaky(x[n, cin] · w[cout, cin]^T + bias, slope); bias may be null; slope = 1: no activation
int linear_forward_device(const float* x, int n, int cin, const float* w, const float* bias, int cout, float slope,
                          float* y, cudaStream_t stream) {
    int rc = check_linear(n, cin, cout, slope);
    if (rc != KP_OK || n == 0) return rc;
    Scratch S(stream);
    const float* xa;
    int cin_p;
    if ((rc = padded_cols(S, x, n, cin, &xa, &cin_p, stream)) != KP_OK) return rc;
    return run_dense("lin_fwd", S, n, xa, cin_p, cin, nullptr, 1.f, w, 1, cin, cout, bias, slope, y, cout, stream);
}

// dx[n, cin] = g · w, dw[cout, cin] = g^T · x with g = dy ⊙ leaky'(y) (y = the forward OUTPUT, or null when the layer
// has no activation); dx may be null. dw is overwritten.
int linear_backward_device(const float* x, int n, int cin, const float* w, int cout, const float* y, float slope,
                           const float* dy, float* dx, float* dw, cudaStream_t stream) {
    int rc = check_linear(n, cin, cout, slope);
    if (rc != KP_OK) return rc;
    KP_CUDA(cudaMemsetAsync(dw, 0, (size_t)cin * cout * sizeof(float), stream));
    if (n == 0) return KP_OK;
    Scratch S(stream);
    const float *ga, *ya = y;
    int cout_p;
    if ((rc = padded_cols(S, dy, n, cout, &ga, &cout_p, stream)) != KP_OK) return rc;
    if (y && cout_p != cout && (rc = padded_cols(S, y, n, cout, &ya, &cout_p, stream)) != KP_OK) return rc;
    if (dx) {
        rc = run_dense("lin_dx", S, n, ga, cout_p, cout, ya, slope, w, cin, 1, cin, nullptr, 1.f, dx, cin, stream);
        if (rc != KP_OK) return rc;
    }
    // dW[o, c] = sum_i g[i, o] x[i, c]: kp_dw with A = g (M = o), B tile = x rows (N = c)
    DwParams P;
    P.nq = n; P.x = ga; P.cin = cout; P.cin_p = cout_p; P.K = 1; P.H = 0;
    P.koff = nullptr; P.entries = nullptr;
    P.dout = x; P.cout = cin;
    const int cin_p16 = (cin + 15) & ~15;
    P.NB = cin_p16 < 256 ? cin_p16 : 256;
    const int n_slices = ceil_div(cin_p16, P.NB);
    const int n_chunks = ceil_div(cout_p, CK);
    P.n_tiles = ceil_div(n, TILE_M);
    const size_t smem = (size_t)4 * MN_LBO + (size_t)dw_b_bytes(P.NB) + TILE_M * KOFF * 2 + TILE_M * 4 + 64;
    int splits = (smem <= (size_t)SMEM_TWO_CTAS ? 2 * 148 : 148) / (n_chunks * n_slices);  // one wave of CTAs
    if (splits < 1) splits = 1;
    if (splits > P.n_tiles) splits = P.n_tiles;
    P.n_splits = splits;
    P.dw = dw;
    P.mask = ya; P.slope_in = slope;
    uint32_t cols = 32;
    while ((int)cols < P.NB) cols <<= 1;
    P.tmem_cols = cols;
    if (smem <= (size_t)SMEM_TWO_CTAS) {
        KP_CUDA(set_smem(kp_dw_kernel<8, true>, smem));
        ProfileScope ps("lin_dw", stream);
        kp_dw_kernel<8, true><<<dim3(n_chunks, splits, n_slices), 256, smem, stream>>>(P);
    } else {
        KP_CUDA(set_smem(kp_dw_kernel<16, true>, smem));
        ProfileScope ps("lin_dw", stream);
        kp_dw_kernel<16, true><<<dim3(n_chunks, splits, n_slices), 512, smem, stream>>>(P);
    }
    KP_CHECK_LAUNCH();
    return KP_OK;
}

}  // namespace kp
