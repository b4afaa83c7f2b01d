// prep_kernels.cu -- index structures of the observations built on the device (they used to be host loops over all
// image points in prepare(): 10^7 entries at config 5): the image of every observation, and the observations of
// every object point (CSC) by a STABLE radix sort of (object point, observation index) -- stable, so every point's
// observations stay in observation order and the by-point sums keep their fixed order.
#include <cub/cub.cuh>

#include "common.h"

namespace jaicov {

__global__ void __launch_bounds__(256) k_img_of_obs(const int64_t *__restrict__ pt_ptr, int32_t *__restrict__ img_of_obs) {
    const int img = blockIdx.x;
    for (int64_t j = pt_ptr[img] + threadIdx.x; j < pt_ptr[img + 1]; j += blockDim.x) img_of_obs[j] = img;
}

void launch_img_of_obs(const int64_t *pt_ptr, int nImg, int32_t *img_of_obs, cudaStream_t s) {
    if (nImg <= 0) return;
    g_launch_count++;
    k_img_of_obs<<<nImg, 256, 0, s>>>(pt_ptr, img_of_obs);
}

__global__ void __launch_bounds__(256) k_iota64(int64_t *__restrict__ v, int64_t n, int64_t first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = first + i;
}

// pt_obs_ptr[p] = first position in the sorted keys with key >= p  (p = 0 .. nPt)
__global__ void __launch_bounds__(256) k_csc_ptr(const int32_t *__restrict__ keys, int64_t n, int nPt, int64_t *__restrict__ ptr) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nPt) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < p) lo = mid + 1; else hi = mid;
    }
    ptr[p] = lo;
}

size_t csc_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int64_t *)nullptr,
                                    (int64_t *)nullptr, n);
    return bytes;
}

// obj = object point of the observations [obs0, obs0 + n); keys_out (n), iota (n) and temp are scratch.
// minmax_host[0/1] receive the smallest / largest object point index (validation by the caller).
void launch_build_csc(const int32_t *obj, int64_t obs0, int64_t n, int nPt, int32_t *keys_out, int64_t *iota, void *temp,
                      size_t temp_bytes, int64_t *pt_obs_ptr, int64_t *pt_obs, int32_t minmax_host[2], cudaStream_t s) {
    minmax_host[0] = 0;
    minmax_host[1] = -1;
    if (n > 0) {
        g_launch_count += 2;
        k_iota64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(iota, n, obs0);
        JCHECK(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, obj + obs0, keys_out, iota, pt_obs, n, 0, 32, s));
        JCHECK(cudaMemcpyAsync(&minmax_host[0], keys_out, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        JCHECK(cudaMemcpyAsync(&minmax_host[1], keys_out + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    g_launch_count++;
    k_csc_ptr<<<(unsigned)((nPt + 1 + 255) / 256), 256, 0, s>>>(keys_out, n, nPt, pt_obs_ptr);
    JCHECK(cudaStreamSynchronize(s));
}

}  // namespace jaicov
