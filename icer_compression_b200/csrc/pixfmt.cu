// pixfmt.cu -- the steps the reference's tools run on the host right before / after the codec, as device kernels
// (SURVEY 8(f) rank 1): 8-bit samples widened into the uint16 containers the *_uint16 API codes
// (example/src/icer_util.c:163-168), packed RGB888 <-> planar Y,Cb,Cr with the tools' integer conversion
// (example/inc/color_util.h:27-34, example/src/icer_util.c:69-121) and the 8-bit display clamp (icer_util.c:322-325).
// They let 8-bit / RGB batches cross PCIe and sit in HBM at 1-3 bytes per pixel instead of 2-6.
#include "common.h"

namespace {
__device__ __forceinline__ int clip255(int x) { return x > 255 ? 255 : (x < 0 ? 0 : x); }

__global__ void __launch_bounds__(256) k_u8_to_u16(const uint8_t *__restrict__ src, uint16_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) k_u16_to_u8(const uint16_t *__restrict__ src, uint8_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
        const uint16_t v = src[i];
        dst[i] = v > 255 ? 255 : (uint8_t) v;                                  // icer_util.c:322-325
    }
}
// native uint8 codec (icer_compress_image_uint8 & co): samples are int8; they travel sign-extended in the int16 containers
__global__ void __launch_bounds__(256) k_s8_to_i16(const uint8_t *__restrict__ src, uint16_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x)
        dst[i] = (uint16_t) (int16_t) (int8_t) src[i];
}
__global__ void __launch_bounds__(256) k_i16_to_b8(const uint16_t *__restrict__ src, uint8_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) dst[i] = (uint8_t) src[i];
}
// sign bit 15 + 7-bit magnitude -> the uint8 codec's sign bit 7 + magnitude (what it leaves in the caller's image buffer)
__global__ void __launch_bounds__(256) k_sm16_to_sm8(const uint16_t *__restrict__ src, uint8_t *__restrict__ dst, size_t count) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
        const uint16_t v = src[i];
        dst[i] = (uint8_t) ((v & 0x7F) | ((v >> 8) & 0x80));
    }
}
// rgb: [n][px][3] packed; yuv: [n][3][px]
__global__ void __launch_bounds__(256) k_rgb_to_yuv(const uint8_t *__restrict__ rgb, uint16_t *__restrict__ yuv, size_t n, size_t px) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n * px; i += (size_t) gridDim.x * blockDim.x) {
        const size_t f = i / px, p = i % px;
        const int r = rgb[i * 3], g = rgb[i * 3 + 1], b = rgb[i * 3 + 2];
        const int y = clip255((19595 * r + 38470 * g + 7471 * b) >> 16);          // CRGB2Y  (color_util.h:27)
        const int cb = clip255(((36962 * (b - y)) >> 16) + 128);                  // CRGB2Cb (:28)
        const int cr = clip255(((46727 * (r - y)) >> 16) + 128);                  // CRGB2Cr (:29)
        uint16_t *o = yuv + f * 3 * px + p;
        o[0] = (uint16_t) y; o[px] = (uint16_t) cb; o[2 * px] = (uint16_t) cr;
    }
}
// yuv: [n][3][cap_px] with the image in the first px entries of every plane; rgb: [n][cap_px][3]
__global__ void __launch_bounds__(256) k_yuv_to_rgb(const uint16_t *__restrict__ yuv, uint8_t *__restrict__ rgb, size_t n, size_t px,
                                                    size_t cap_px) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n * px; i += (size_t) gridDim.x * blockDim.x) {
        const size_t f = i / px, p = i % px;
        const uint16_t *s = yuv + f * 3 * cap_px + p;
        const int y = s[0], cb = s[cap_px], cr = s[2 * cap_px];
        uint8_t *o = rgb + (f * cap_px + p) * 3;
        o[0] = (uint8_t) clip255(y + ((91881 * cr) >> 16) - 179);                 // CYCbCr2R (color_util.h:32)
        o[1] = (uint8_t) clip255(y - ((22544 * cb + 46793 * cr) >> 16) + 135);    // CYCbCr2G (:33)
        o[2] = (uint8_t) clip255(y + ((116129 * cb) >> 16) - 226);                // CYCbCr2B (:34)
    }
}
unsigned grid_for(size_t count) {
    size_t b = (count + 255) / 256;
    return (unsigned) (b > 148 * 32 ? 148 * 32 : (b ? b : 1));
}
}  // namespace

int launch_u8_to_u16(const uint8_t *src, uint16_t *dst, size_t count, cudaStream_t st, LaunchStats &ls) {
    k_u8_to_u16<<<grid_for(count), 256, 0, st>>>(src, dst, count);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_u16_to_u8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls) {
    k_u16_to_u8<<<grid_for(count), 256, 0, st>>>(src, dst, count);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_s8_to_i16(const uint8_t *src, uint16_t *dst, size_t count, cudaStream_t st, LaunchStats &ls) {
    k_s8_to_i16<<<grid_for(count), 256, 0, st>>>(src, dst, count);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_i16_to_b8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls) {
    k_i16_to_b8<<<grid_for(count), 256, 0, st>>>(src, dst, count);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_sm16_to_sm8(const uint16_t *src, uint8_t *dst, size_t count, cudaStream_t st, LaunchStats &ls) {
    k_sm16_to_sm8<<<grid_for(count), 256, 0, st>>>(src, dst, count);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_rgb_to_yuv(const uint8_t *rgb, uint16_t *yuv, size_t n, size_t px, cudaStream_t st, LaunchStats &ls) {
    k_rgb_to_yuv<<<grid_for(n * px), 256, 0, st>>>(rgb, yuv, n, px);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
int launch_yuv_to_rgb(const uint16_t *yuv, uint8_t *rgb, size_t n, size_t px, size_t cap_px, cudaStream_t st, LaunchStats &ls) {
    k_yuv_to_rgb<<<grid_for(n * px), 256, 0, st>>>(yuv, rgb, n, px, cap_px);
    ls.launches++;
    CUDA_TRY(cudaGetLastError());
    return ICER_RESULT_OK;
}
