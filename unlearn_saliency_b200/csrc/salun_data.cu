// salun_data.cu -- the steps either side of the hot path as kernels:
//
//   salun_augment_batch  the reference's training transform RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor
//                        (Classification/dataset.py:549-555) plus the batch gather of the DataLoader, from a uint8 dataset
//                        RESIDENT in HBM (CIFAR-10 train: 154 MB) straight into the fp32 NCHW batch the engine consumes.
//                        At 400+ steps/s (100 k images/s) the reference's num_workers=0 PIL pipeline is the wall
//                        (main_forget.py:42-48); this kernel moves 3 KB in and 12 KB out per image.
//   salun_eval_logits    what trainer/val.py:44-61 and evaluation/SVC_MIA.py:44-46 do with the logits of one batch:
//                        cross-entropy (summed), top-1 hits, and optionally the softmax probabilities -- accumulated on the
//                        device so that a validation pass needs ONE device->host read instead of two .item() per batch.
#include <math.h>

#include "salun_common.cuh"

namespace salun {

// one thread per output element (n, c, y, x): out = img[index[n]][y + dy - pad][xs + dx - pad][c] / 255, xs = flip ? W-1-x : x
__global__ void __launch_bounds__(256) k_augment(const uint8_t *__restrict__ img, const int64_t *__restrict__ index,
                                                 const int32_t *__restrict__ crop_xy, const uint8_t *__restrict__ flip,
                                                 long long total, int H, int W, int pad, long long n_images,
                                                 float *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % 3);
    const long long n = i / ((long long)W * H * 3);
    long long src = index ? index[n] : n;
    src = src < 0 ? 0 : (src >= n_images ? n_images - 1 : src);  // no out-of-dataset reads on a bad index
    const int dx = crop_xy ? crop_xy[2 * n] : pad, dy = crop_xy ? crop_xy[2 * n + 1] : pad;
    const int xs = (flip && flip[n]) ? W - 1 - x : x;   // the flip acts on the cropped image (transform order)
    const int sy = y + dy - pad, sx = xs + dx - pad;
    float v = 0.f;                                      // RandomCrop pads with zeros
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = (float)img[((src * H + sy) * W + sx) * 3 + c];
    out[i] = __fdiv_rn(v, 255.f);                       // ToTensor: uint8 -> float32 .div(255)
  }
}

// one block per launch: rows strided over threads, fixed-order block reduction, then ONE thread updates the accumulators
__global__ void __launch_bounds__(256) k_eval_logits(const float *__restrict__ logits, const int64_t *__restrict__ labels, int n,
                                                     int K, float *__restrict__ probs, double *__restrict__ loss_sum,
                                                     long long *__restrict__ correct) {
  __shared__ double sl[256];
  __shared__ int sc[256];
  double l = 0.0;
  int c = 0;
  for (int r = threadIdx.x; r < n; r += 256) {
    const float *row = logits + (size_t)r * K;
    float mx = row[0];
    int arg = 0;
    for (int k = 1; k < K; ++k)
      if (row[k] > mx) {   // first maximum wins, like torch.argmax / topk on ties
        mx = row[k];
        arg = k;
      }
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(row[k] - mx);
    if (probs) {
      const float inv = 1.f / se;
      for (int k = 0; k < K; ++k) probs[(size_t)r * K + k] = expf(row[k] - mx) * inv;
    }
    if (labels) {
      long long y = labels[r];
      y = y < 0 ? 0 : (y >= K ? K - 1 : y);
      l += (double)((logf(se) + mx) - row[y]);
      c += arg == (int)y;
    }
  }
  sl[threadIdx.x] = l;
  sc[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0 && labels) {
    double tl = 0.0;
    long long tc = 0;
    for (int i = 0; i < 256; ++i) {
      tl += sl[i];
      tc += sc[i];
    }
    if (loss_sum) *loss_sum += tl;
    if (correct) *correct += tc;
  }
}

}  // namespace salun

using namespace salun;

extern "C" {

int salun_augment_batch(salun_ctx *ctx, const uint8_t *images_hwc, int64_t n_images, const int64_t *index,
                        const int32_t *crop_xy, const uint8_t *flip, int n, int H, int W, int pad, float *out_nchw,
                        void *stream) {
  SALUN_REQUIRE(ctx && images_hwc && out_nchw, "NULL argument");
  SALUN_REQUIRE(n >= 0 && H > 0 && W > 0 && pad >= 0 && n_images > 0, "bad sizes");
  if (n == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)n * 3 * H * W;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  k_augment<<<(int)g, 256, 0, (cudaStream_t)stream>>>(images_hwc, index, crop_xy, flip, total, H, W, pad, n_images, out_nchw);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

int salun_eval_logits(salun_ctx *ctx, const float *logits, const int64_t *labels, int n, int K, float *probs,
                      double *loss_sum_dev, int64_t *correct_dev, void *stream) {
  SALUN_REQUIRE(ctx && logits, "NULL argument");
  SALUN_REQUIRE(n >= 0 && K > 0, "bad sizes");
  SALUN_REQUIRE(!labels || loss_sum_dev || correct_dev, "labels given but no accumulator");
  if (n == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  k_eval_logits<<<1, 256, 0, (cudaStream_t)stream>>>(logits, labels, n, K, probs, loss_sum_dev, (long long *)correct_dev);
  ++g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// One step of the conditional DDIM / generalized sampler (DDPM/functions/denoising.py:72-95) after the two U-Net passes:
//   et      = (1 + s) * eps_cond - s * eps_null                      (classifier-free guidance, models/diffusion.py:340-355)
//   x0_t    = (xt - et * sqrt(1 - at)) / sqrt(at)
//   c1      = eta * sqrt((1 - at / at_next) * (1 - at_next) / (1 - at)) ;  c2 = sqrt((1 - at_next) - c1^2)
//   xt_next = sqrt(at_next) * x0_t + c1 * noise + c2 * et
// per-sample at / at_next (compute_alpha, :4-7); the same fp32 operation order as the torch statements.
// ---------------------------------------------------------------------------------------------------------------------
namespace salun {
__global__ void __launch_bounds__(256) k_ddim_step(const float *__restrict__ eps_cond, const float *__restrict__ eps_null,
                                                   const float *__restrict__ xt, const float *__restrict__ noise,
                                                   const float *__restrict__ at, const float *__restrict__ at_next, float s,
                                                   float eta, long long total, int chw, float *__restrict__ x_next,
                                                   float *__restrict__ x0_out) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int n = (int)(i / chw);
    const float a = at[n], an = at_next[n];
    float et = eps_cond[i];
    if (eps_null) et = __fsub_rn(__fmul_rn(1.f + s, et), __fmul_rn(s, eps_null[i]));
    const float x0 = __fdiv_rn(__fsub_rn(xt[i], __fmul_rn(et, sqrtf(1.f - a))), sqrtf(a));
    const float c1 = eta * sqrtf(__fdiv_rn(__fmul_rn(1.f - __fdiv_rn(a, an), 1.f - an), 1.f - a));
    const float c2 = sqrtf(__fsub_rn(1.f - an, __fmul_rn(c1, c1)));
    float v = __fmul_rn(sqrtf(an), x0);
    v = __fadd_rn(v, __fmul_rn(c1, noise ? noise[i] : 0.f));
    v = __fadd_rn(v, __fmul_rn(c2, et));
    x_next[i] = v;
    if (x0_out) x0_out[i] = x0;
  }
}
}  // namespace salun

extern "C" int salun_ddim_step(salun_ctx *ctx, const float *eps_cond, const float *eps_null, const float *xt, const float *noise,
                               const float *at, const float *at_next, float cond_scale, float eta, int n, int chw,
                               float *x_next, float *x0_out, void *stream) {
  SALUN_REQUIRE(ctx && eps_cond && xt && at && at_next && x_next, "NULL argument");
  SALUN_REQUIRE(n >= 0 && chw > 0, "bad sizes");
  SALUN_REQUIRE(eta == 0.f || noise, "eta != 0 needs the noise tensor");
  if (n == 0) return SALUN_OK;
  SALUN_CUDA_OK(cudaSetDevice(ctx->device));
  const long long total = (long long)n * chw;
  long long g = (total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  salun::k_ddim_step<<<(int)g, 256, 0, (cudaStream_t)stream>>>(eps_cond, eps_null, xt, noise, at, at_next, cond_scale, eta, total,
                                                               chw, x_next, x0_out);
  ++salun::g_launch_count;
  SALUN_CUDA_OK(cudaGetLastError());
  return SALUN_OK;
}
