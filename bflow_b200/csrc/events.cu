// Rows (f1)/(f2) of the scope table: the steps immediately before and after RAFTSpline.forward.
//
//   bflow_voxelize        events -> voxel grid, VoxelGrid.convert (data/utils/representations.py:64-111): bilinear in time (and in
//                         x, y for sub-pixel coordinates) scatter-add.  The reference does this on the CPU with Tensor.put_(accumulate=True)
//                         in DataLoader workers; here it is one atomic-add kernel (float atomics: the summation order differs from
//                         the sequential CPU loop, results agree to rounding).
//   bflow_voxel_norm      norm_voxel_grid (representations.py:9-18): mean / unbiased std over the non-zero voxels, applied to them.
//   bflow_epe_masked      epe_masked (utils/metrics.py:196-213): sum over valid pixels of sqrt(sum_c (src-tgt)^2) and their count — the
//                         per-rank state that bflow_b200.dist.gather_epe exchanges.
//   bflow_flow_metrics    EPE, angular error (ae_masked, metrics.py:259-296) and N-pixel error (n_pixel_error_masked, :161-193) in one
//                         pass, with an optional scale on the prediction (linear-assumption baseline, :298-305).
#include "common.cuh"

namespace bflow {

template <bool FLOAT_XY>
__global__ void voxelize_kernel(const void* __restrict__ xs, const void* __restrict__ ys, const unsigned char* __restrict__ pol,
                                const long long* __restrict__ ts, long long n, long long t0c, long long t1c, int C, int H, int W,
                                float* __restrict__ out, int* __restrict__ oob) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // t_norm = (time - t0)/(t1 - t0)*(C-1) in fp32, as torch computes it for an int64 tensor divided by a Python int
        const float tn = (float)(ts[i] - t0c) / (float)(t1c - t0c) * (float)(C - 1);
        const int tf = (int)floorf(tn);
        const float value = 2.f * (float)pol[i] - 1.f;
        if (!FLOAT_XY) {
            const long long x = reinterpret_cast<const long long*>(xs)[i], y = reinterpret_cast<const long long*>(ys)[i];
            // the reference's put_ raises on an index outside the grid; here the event is dropped and reported, never written
            if (x < 0 || x >= W || y < 0 || y >= H) {
                if (oob != nullptr) atomicAdd(oob, 1);
                continue;
            }
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) {
                const int tl = tf + dt;
                if (tl >= 0 && tl < C) atomicAdd(out + ((size_t)tl * H + y) * W + x, value * (1.f - fabsf((float)tl - tn)));
            }
        } else {
            const float x = reinterpret_cast<const float*>(xs)[i], y = reinterpret_cast<const float*>(ys)[i];
            const int x0 = (int)floorf(x), y0 = (int)floorf(y);
#pragma unroll
            for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dt = 0; dt < 2; ++dt) {
                        const int xl = x0 + dx, yl = y0 + dy, tl = tf + dt;
                        if (xl < W && xl >= 0 && yl < H && yl >= 0 && tl >= 0 && tl < C) {
                            const float wgt = value * (1.f - fabsf((float)xl - x)) * (1.f - fabsf((float)yl - y)) * (1.f - fabsf((float)tl - tn));
                            atomicAdd(out + ((size_t)tl * H + yl) * W + xl, wgt);
                        }
                    }
        }
    }
}

// stats[0] = sum, stats[1] = sum of squares, stats[2] = count over the non-zero entries
__global__ void voxel_nonzero_stats_kernel(const float* __restrict__ v, long long n, double* __restrict__ stats) {
    double s = 0.0, ss = 0.0, c = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = v[i];
        if (x != 0.f) { s += x; ss += (double)x * x; c += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats, s);
        atomicAdd(stats + 1, ss);
        atomicAdd(stats + 2, c);
    }
}

__global__ void voxel_norm_apply_kernel(float* __restrict__ v, long long n, const double* __restrict__ stats) {
    const double cnt = stats[2];
    if (cnt <= 0.0) return;
    const double mean = stats[0] / cnt;
    // torch.std: unbiased (N-1); a single element gives nan in torch -> the "std > 0" test fails and only the mean is removed
    double var = cnt > 1.0 ? (stats[1] - cnt * mean * mean) / (cnt - 1.0) : 0.0;
    if (var < 0.0) var = 0.0;
    const float m = (float)mean, sd = (float)sqrt(var);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = v[i];
        if (x != 0.f) v[i] = sd > 0.f ? (x - m) / sd : (x - m);
    }
}

__global__ void epe_masked_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const unsigned char* __restrict__ valid, int N,
                                  int Cc, long long HW, double* __restrict__ out) {
    double s = 0.0, c = 0.0;
    const long long total = (long long)N * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (valid != nullptr && valid[i] == 0) continue;
        const long long n = i / HW, p = i - n * HW;
        float acc = 0.f;
        for (int ch = 0; ch < Cc; ++ch) {
            const float dlt = src[((size_t)n * Cc + ch) * HW + p] - tgt[((size_t)n * Cc + ch) * HW + p];
            acc = fmaf(dlt, dlt, acc);
        }
        s += (double)sqrtf(acc);
        c += 1.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, s);
        atomicAdd(out + 1, c);
    }
}

// one pass: EPE sum, valid count, angular-error sum (radians), N-pixel-error counts for up to 4 thresholds
struct MetricThresholds {
    float t[4];
    int n;
};
__global__ void flow_metrics_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const unsigned char* __restrict__ valid, int N,
                                    int Cc, long long HW, float src_scale, MetricThresholds th, double* __restrict__ out) {
    double s = 0.0, c = 0.0, ae = 0.0;
    double npe[4] = {0.0, 0.0, 0.0, 0.0};
    const long long total = (long long)N * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (valid != nullptr && valid[i] == 0) continue;
        const long long n = i / HW, p = i - n * HW;
        float e2 = 0.f, dot = 1.f, ns = 1.f, nt = 1.f, gt2 = 0.f;       // the "+1": homogeneous extension of both vectors (metrics.py:266-272)
        for (int ch = 0; ch < Cc; ++ch) {
            const float a = src_scale * src[((size_t)n * Cc + ch) * HW + p], b = tgt[((size_t)n * Cc + ch) * HW + p];
            const float dlt = a - b;
            e2 = fmaf(dlt, dlt, e2);
            dot = fmaf(a, b, dot);
            ns = fmaf(a, a, ns);
            nt = fmaf(b, b, nt);
            gt2 = fmaf(b, b, gt2);
        }
        const float e = sqrtf(e2);
        s += (double)e;
        c += 1.0;
        const float cs = fminf(fmaxf(dot / (sqrtf(ns) * sqrtf(nt)), -1.f), 1.f);
        ae += (double)acosf(cs);
        const float rel = e / fmaxf(sqrtf(gt2), 1e-6f);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < th.n && e > th.t[k] && rel >= 0.05f) npe[k] += 1.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        ae += __shfl_xor_sync(0xffffffffu, ae, o);
#pragma unroll
        for (int k = 0; k < 4; ++k) npe[k] += __shfl_xor_sync(0xffffffffu, npe[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, s);
        atomicAdd(out + 1, c);
        atomicAdd(out + 2, ae);
        for (int k = 0; k < th.n; ++k) atomicAdd(out + 3 + k, npe[k]);
    }
}

static inline unsigned grid_for(long long total) {
    long long g = (total + 255) / 256;
    if (g < 1) g = 1;
    const long long cap = (long long)bflow::num_sms() * 16;
    if (g > cap) g = cap;
    return (unsigned)g;
}

}  // namespace bflow

extern "C" int bflow_voxelize(const void* x, const void* y, int xy_is_float, const unsigned char* pol, const long long* time, long long n_events,
                              long long t0_center, long long t1_center, int channels, int H, int W, float* out, int* oob_count, void* stream) {
    BFLOW_REQUIRE(out != nullptr && channels > 1 && H > 1 && W > 1, "voxelize: bad grid (representations.py:28-34)");
    BFLOW_REQUIRE(n_events >= 0 && (n_events == 0 || (x != nullptr && y != nullptr && pol != nullptr && time != nullptr)), "voxelize: null events");
    BFLOW_REQUIRE(t1_center > t0_center, "voxelize: t1_center must be greater than t0_center");
    if (n_events == 0) return BFLOW_OK;
    if (xy_is_float)
        bflow::voxelize_kernel<true><<<bflow::grid_for(n_events), 256, 0, (cudaStream_t)stream>>>(x, y, pol, time, n_events, t0_center, t1_center, channels, H, W, out, oob_count);
    else
        bflow::voxelize_kernel<false><<<bflow::grid_for(n_events), 256, 0, (cudaStream_t)stream>>>(x, y, pol, time, n_events, t0_center, t1_center, channels, H, W, out, oob_count);
    return bflow::check_launch("bflow_voxelize");
}

extern "C" int bflow_voxel_norm(float* voxel, long long numel, double* stats3, void* stream) {
    BFLOW_REQUIRE(voxel != nullptr && stats3 != nullptr && numel > 0, "voxel_norm: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(stats3, 0, 3 * sizeof(double), st);
    bflow::voxel_nonzero_stats_kernel<<<bflow::grid_for(numel), 256, 0, st>>>(voxel, numel, stats3);
    bflow::voxel_norm_apply_kernel<<<bflow::grid_for(numel), 256, 0, st>>>(voxel, numel, stats3);
    return bflow::check_launch("bflow_voxel_norm");
}

extern "C" int bflow_epe_masked(const float* src, const float* tgt, const unsigned char* valid, int N, int C, long long HW, double* sum_count,
                                void* stream) {
    BFLOW_REQUIRE(src != nullptr && tgt != nullptr && sum_count != nullptr && N > 0 && C > 0 && HW > 0, "epe_masked: bad arguments");
    bflow::epe_masked_kernel<<<bflow::grid_for((long long)N * HW), 256, 0, (cudaStream_t)stream>>>(src, tgt, valid, N, C, HW, sum_count);
    return bflow::check_launch("bflow_epe_masked");
}

extern "C" int bflow_flow_metrics(const float* src, const float* tgt, const unsigned char* valid, int N, int C, long long HW, float src_scale,
                                  const float* thresholds_host, int n_thresholds, double* out8, void* stream) {
    BFLOW_REQUIRE(src != nullptr && tgt != nullptr && out8 != nullptr && N > 0 && C > 0 && HW > 0, "flow_metrics: bad arguments");
    BFLOW_REQUIRE(n_thresholds >= 0 && n_thresholds <= 4 && (n_thresholds == 0 || thresholds_host != nullptr), "flow_metrics: at most 4 N-pixel thresholds");
    bflow::MetricThresholds th{};
    th.n = n_thresholds;
    for (int k = 0; k < n_thresholds; ++k) {
        BFLOW_REQUIRE(thresholds_host[k] > 0.f, "flow_metrics: n_pixels must be positive (metrics.py:140)");
        th.t[k] = thresholds_host[k];
    }
    bflow::flow_metrics_kernel<<<bflow::grid_for((long long)N * HW), 256, 0, (cudaStream_t)stream>>>(src, tgt, valid, N, C, HW, src_scale, th, out8);
    return bflow::check_launch("bflow_flow_metrics");
}
