// Operand staging for the tensor-core kernels: packs fp32 NHWC rows into the SWIZZLE_128B B-operand image
// (hi | lo fp16 planes per k-block) that conv_tc3.cu loads with one cp.async.bulk per tile.  Used for the all-pairs correlation
// GEMM (models/raft_utils/corr.py:264-272), whose "weights" are the target feature map.
#include "common.cuh"

namespace bflow {
// fp32 NHWC rows -> tensor-core B-operand image (see bflow_pack_b_tc in the header)
__global__ void pack_b_tc_kernel(const float* __restrict__ src, int ld, uint8_t* __restrict__ dst, int rows, int K, int bn, int nkb,
                                 int plane_h, int plane_w) {
    const int chunks = (K + 7) / 8;
    const long long total = (long long)rows * chunks;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / chunks);
        const int ck = (int)(idx - (long long)r * chunks);
        const int k = ck * 8;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = (k + e < K) ? __ldg(src + (size_t)r * ld + k + e) : 0.f;
        uint4 hi, lo;
        split2(x[0], x[1], hi.x, lo.x);
        split2(x[2], x[3], hi.y, lo.y);
        split2(x[4], x[5], hi.z, lo.z);
        split2(x[6], x[7], hi.w, lo.w);
        int n = r;
        if (plane_w > 0) {               // source rows are pixels (y, x): land in 4x4-tiled order
            const int y = r / plane_w, xx = r - y * plane_w;
            const int wp4 = (plane_w + 3) >> 2;
            n = (((y >> 2) * wp4 + (xx >> 2)) << 4) + ((y & 3) << 2) + (xx & 3);
        }
        const int tile = n / bn, rr = n - tile * bn;
        const int kb = ck >> 3, c = ck & 7;
        uint8_t* base = dst + ((size_t)tile * nkb + kb) * (size_t)(2 * bn * 128) + (size_t)rr * 128 + (size_t)((c ^ (rr & 7)) << 4);
        *reinterpret_cast<uint4*>(base) = hi;
        *reinterpret_cast<uint4*>(base + (size_t)bn * 128) = lo;
    }
}
}  // namespace bflow

extern "C" int bflow_pack_b_tc(const float* src, int ld, void* dst, int rows, int K, int bn, int plane_h, int plane_w, void* stream) {
    BFLOW_REQUIRE(src != nullptr && dst != nullptr, "pack_b_tc: null tensor");
    BFLOW_REQUIRE(rows > 0 && K > 0 && ld >= K && (bn == 64 || bn == 128 || bn == 256), "pack_b_tc: bad shape");
    BFLOW_REQUIRE(plane_w == 0 || (plane_h > 0 && plane_h * plane_w == rows), "pack_b_tc: plane does not match rows");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0, "pack_b_tc: dst must be 16-byte aligned");
    const int nkb = (K + 63) / 64;
    const long long total = (long long)rows * ((K + 7) / 8);
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    bflow::pack_b_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld, reinterpret_cast<uint8_t*>(dst), rows, K, bn, nkb, plane_h, plane_w);
    return bflow::check_launch("bflow_pack_b_tc");
}
