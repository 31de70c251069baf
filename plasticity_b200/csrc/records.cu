// Device-side decoding of labelled byte records — the CIFAR-10 binary format read by
// nnet/cifar_test.cc:191-205: one label byte followed by `sample_size` pixel bytes —
// into what the network trains on:
//   input   x[i] = (double)(signed char)pixel[i] / ||pixels||_2   (norm 0 -> 1),
//           Sample::NormalizedInput, cifar_test.cc:66-81 (the reference reads the bytes
//           as signed `char`, :55-62)
//   target  one-hot of the label, Sample::OneHotEncodedOutput, cifar_test.cc:83-97
// so that a host-resident data set crosses PCIe as bytes (3 073 B per image instead of
// 12 328 B of fp32) and is normalised where it is consumed.  The sum of squares is exact
// (integers), the quotient is formed in double and rounded to fp32 once: bit-identical to
// normalising in double on the host and converting to float.  A pixel has 256 possible
// values, so the 256 quotients of a record are formed once (one per thread) and the pixels
// look theirs up: 256 double divisions per image instead of 3 072 — the kernel shares the
// GPU with the training step of the previous batch and should stay out of its way (with the
// pixels fetched once as words instead of twice as bytes the end-to-end rate went from 4.32 M
// to 4.44 M samples/s, at 4.59 M device-resident).
#include "common.cuh"
#include "kernels.cuh"

namespace pb {

// One CTA per record.  The record's pixels are fetched ONCE, as aligned 32-bit words, into
// shared memory (records are 3 073 bytes long, so their pixels start at any byte offset); both
// passes - the sum of squares and the table look-up - read them from there, and the outputs
// leave as float4.
__global__ void __launch_bounds__(256)
decode_records_kernel(const uint8_t *__restrict__ rec, float *__restrict__ x,
                      float *__restrict__ e, int sample_size, int n_classes, int64_t total_bytes) {
  extern __shared__ uint32_t raw[];
  __shared__ unsigned long long red[8];
  __shared__ double norm_s;
  __shared__ float quot[256];  // quot[v + 128] = (float)((double)v / norm)
  const uint8_t *r = rec + (int64_t)blockIdx.x * (sample_size + 1);
  const uint8_t *rec_end = rec + total_bytes;
  const uintptr_t p0 = (uintptr_t)(r + 1);
  const int off = (int)(p0 & 3);
  const uint8_t *w0 = reinterpret_cast<const uint8_t *>(p0 - off);
  const int nwords = (off + sample_size + 3) >> 2;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
    const uint8_t *q = w0 + 4 * i;
    uint32_t v = 0;
    if (q >= rec && q + 4 <= rec_end) {
      v = __ldg(reinterpret_cast<const uint32_t *>(q));
    } else {  // the word straddles an end of the buffer: only the bytes inside it
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (q + k >= rec && q + k < rec_end) v |= (uint32_t)q[k] << (8 * k);
    }
    raw[i] = v;
  }
  __syncthreads();
  const signed char *px = reinterpret_cast<const signed char *>(raw) + off;
  unsigned long long sum = 0;
  for (int i = threadIdx.x; i < sample_size; i += blockDim.x) {
    const int v = px[i];
    sum += (unsigned long long)(v * v);
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    double norm = sqrt((double)t);
    if (norm == 0) norm = 1;
    norm_s = norm;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < 256; v += blockDim.x)
    quot[v] = (float)((double)(v - 128) / norm_s);
  __syncthreads();
  float *xo = x + (int64_t)blockIdx.x * sample_size;
  if ((sample_size & 3) == 0 && (((uintptr_t)x) & 15) == 0) {
    for (int i = threadIdx.x; i < sample_size / 4; i += blockDim.x)
      reinterpret_cast<float4 *>(xo)[i] =
          make_float4(quot[(int)px[4 * i] + 128], quot[(int)px[4 * i + 1] + 128],
                      quot[(int)px[4 * i + 2] + 128], quot[(int)px[4 * i + 3] + 128]);
  } else {
    for (int i = threadIdx.x; i < sample_size; i += blockDim.x) xo[i] = quot[(int)px[i] + 128];
  }
  const int label = r[0];
  float *eo = e + (int64_t)blockIdx.x * n_classes;
  for (int i = threadIdx.x; i < n_classes; i += blockDim.x) eo[i] = i == label ? 1.0f : 0.0f;
}

}  // namespace pb

namespace pb {

// On `stream` (the host pipeline decodes on its copy stream, right behind the H2D copy,
// so the decode of step i+1 overlaps the compute of step i).
int records_decode_on(pb_context *ctx, cudaStream_t stream, const uint8_t *d_records,
                      float *d_inputs, float *d_expected, int64_t n_records, int64_t sample_size,
                      int64_t n_classes) {
  PB_CHECK(ctx && d_records && d_inputs && d_expected, "pb_records_decode: null");
  PB_CHECK(n_records >= 0 && sample_size >= 1 && n_classes >= 1 && n_records < (1ll << 31) &&
               sample_size < (1ll << 31) && n_classes < (1ll << 31),
           "pb_records_decode: bad sizes");
  if (n_records == 0) return 0;
  const size_t smem = (size_t)(sample_size + 8);
  PB_CHECK(smem <= 48 * 1024, "pb_records_decode: records longer than 48 KB are not supported");
  decode_records_kernel<<<(unsigned)n_records, 256, smem, stream>>>(
      d_records, d_inputs, d_expected, (int)sample_size, (int)n_classes,
      n_records * (sample_size + 1));
  PB_LAUNCHED(ctx);
  return 0;
}

}  // namespace pb

extern "C" int pb_records_decode(pb_context *ctx, const uint8_t *d_records, float *d_inputs,
                                 float *d_expected, int64_t n_records, int64_t sample_size,
                                 int64_t n_classes) {
  PB_CHECK(ctx, "pb_records_decode: null");
  return pb::records_decode_on(ctx, ctx->stream, d_records, d_inputs, d_expected, n_records,
                               sample_size, n_classes);
}
