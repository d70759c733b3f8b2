// Kaldi compressed-matrix decode on the device: egs feature payloads -> FP16 rows of the padded minibatch layout.
//
// The reference decodes on the CPU while parsing the egs archive (/root/reference/internal/parser/matrix.go:11-165:
// ReadCompressedMatrix "CM" = one byte per element + per-column percentile headers, column-major bytes;
// ReadCompressedMatrix2 "CM2" = uint16 row-major; ReadCompressedMatrix3 "CM3" = uint8 row-major; ReadFullMatrix "FM" =
// float32), converts to FP16 on the CPU (internal/gpu/bridge.go:141) and uploads FP16.  Here the PAYLOAD BYTES as they
// sit in the archive travel to the device (1-2 bytes per element instead of 2-4) and one launch decodes every sequence
// of the minibatch straight into its FP16 rows: same float32 arithmetic, operation for operation (no FMA contraction),
// then the same round-to-nearest-even FP16 conversion -- the result is bit-identical to the reference's path.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/kaldi_fp16_fused.h"
#include "host_common.h"

using namespace kfp16;

namespace {

constexpr int kMaxMats = 64;       // matrices per launch (kernel-parameter table)
struct DecodeTable {
  int count;
  kfp16_cm_desc d[kMaxMats];
};

// matrix.go:11-14 uint16ToFloat: globalMin + globalRange * (1/65535) * value, evaluated left to right in float32
__device__ __forceinline__ float u16_to_float(float gmin, float grange, unsigned v) {
  const float inv65535 = 1.52590218966964e-05f;
  return __fadd_rn(gmin, __fmul_rn(__fmul_rn(grange, inv65535), (float)v));
}
// matrix.go:17-26 charToFloat
__device__ __forceinline__ float char_to_float(float p0, float p25, float p75, float p100, unsigned v) {
  if (v <= 64) return __fadd_rn(p0, __fmul_rn(__fmul_rn(__fsub_rn(p25, p0), (float)v), 1.0f / 64.0f));
  if (v <= 192) return __fadd_rn(p25, __fmul_rn(__fmul_rn(__fsub_rn(p75, p25), (float)(v - 64)), 1.0f / 128.0f));
  // branch 3: the product in float32, the division and the sum in double (matches Kaldi)
  return (float)__dadd_rn((double)p75, __ddiv_rn((double)__fmul_rn(__fsub_rn(p100, p75), (float)(v - 192)), 63.0));
}

__global__ void decode_kernel(const unsigned char* __restrict__ payload, __half* __restrict__ dst, int ld, DecodeTable tab) {
  const kfp16_cm_desc d = tab.d[blockIdx.y];
  const unsigned char* src = payload + d.payload_offset;
  __half* out = dst + (size_t)d.dst_row * ld;
  const size_t total = (size_t)d.rows * d.cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int r = (int)(i / d.cols), c = (int)(i % d.cols);
    float v;
    switch (d.format) {
      case KFP16_CM: {       // [cols x 4 uint16 percentiles][bytes, COLUMN-major]
        const unsigned short* hdr = reinterpret_cast<const unsigned short*>(src) + (size_t)c * 4;
        const float p0 = u16_to_float(d.global_min, d.global_range, hdr[0]), p25 = u16_to_float(d.global_min, d.global_range, hdr[1]);
        const float p75 = u16_to_float(d.global_min, d.global_range, hdr[2]), p100 = u16_to_float(d.global_min, d.global_range, hdr[3]);
        v = char_to_float(p0, p25, p75, p100, src[(size_t)d.cols * 8 + (size_t)c * d.rows + r]);
        break;
      }
      case KFP16_CM2:        // uint16 row-major: globalMin + value * (globalRange / 65535)
        v = __fadd_rn(d.global_min, __fmul_rn((float)reinterpret_cast<const unsigned short*>(src)[i], __fdiv_rn(d.global_range, 65535.0f)));
        break;
      case KFP16_CM3:        // uint8 row-major: globalMin + value * (globalRange / 255)
        v = __fadd_rn(d.global_min, __fmul_rn((float)src[i], __fdiv_rn(d.global_range, 255.0f)));
        break;
      default:               // KFP16_FM: float32 row-major (possibly unaligned inside the archive)
        { unsigned u = src[4 * i] | (src[4 * i + 1] << 8) | (src[4 * i + 2] << 16) | ((unsigned)src[4 * i + 3] << 24); v = __uint_as_float(u); }
        break;
    }
    out[(size_t)r * ld + c] = __float2half_rn(v);       // fp16.ConvertFloat32ToFloat16: round to nearest even
  }
}

size_t payload_bytes(const kfp16_cm_desc& d) {
  const size_t n = (size_t)d.rows * d.cols;
  switch (d.format) {
    case KFP16_CM: return (size_t)d.cols * 8 + n;
    case KFP16_CM2: return n * 2;
    case KFP16_CM3: return n;
    default: return n * 4;
  }
}

}  // namespace

extern "C" {

size_t kfp16_cm_payload_bytes(const kfp16_cm_desc* d) { return d ? payload_bytes(*d) : 0; }

int kfp16_decode_matrices(kfp16_ctx* ctx, const void* payload_dev, size_t payload_size, const kfp16_cm_desc* descs, int count,
                          void* dst_f16, int ld, int dst_rows) {
  if (count <= 0) return 0;
  if (!payload_dev || !descs || !dst_f16) { set_error("kfp16_decode_matrices: null pointer"); return -1; }
  cudaStream_t s = ctx ? ctx->stream : default_stream();
  for (int base = 0; base < count; base += kMaxMats) {
    DecodeTable tab;
    tab.count = count - base < kMaxMats ? count - base : kMaxMats;
    size_t most = 1;
    for (int i = 0; i < tab.count; ++i) {
      const kfp16_cm_desc& d = descs[base + i];
      if (d.format < KFP16_CM || d.format > KFP16_FM || d.rows < 0 || d.cols < 0 || d.cols > ld || d.dst_row < 0 || d.dst_row + d.rows > dst_rows ||
          d.payload_offset + payload_bytes(d) > payload_size || (d.format != KFP16_CM3 && d.format != KFP16_FM && (d.payload_offset & 1))) {
        set_error("kfp16_decode_matrices: matrix %d (format %d, %d x %d at payload offset %zu, destination row %d) does not fit the payload / destination",
                  base + i, d.format, d.rows, d.cols, (size_t)d.payload_offset, d.dst_row);
        return -1;
      }
      tab.d[i] = d;
      most = std::max(most, (size_t)d.rows * d.cols);
    }
    int gx = (int)std::min<size_t>((most + 255) / 256, 148 * 8 / (size_t)std::max(1, std::min(tab.count, 8)) + 1);
    decode_kernel<<<dim3(gx, tab.count), 256, 0, s>>>((const unsigned char*)payload_dev, (__half*)dst_f16, ld, tab);
    count_launch();
    if (!check_launch("kfp16_decode_matrices")) return -1;
  }
  return 0;
}

}  // extern "C"
