// salun_resnetb.cuh -- internal interface of the Bottleneck-ResNet runtime (salun_resnetb.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "salun_common.cuh"

namespace salun {
struct FlatNet;
int64_t flatnet_param_count(const salun_resnet_cfg *cfg, int64_t *n_bn);
int flatnet_create(salun_ctx *ctx, const salun_resnet_cfg *cfg, float *params, float *grads, float *rmean, float *rvar,
                   FlatNet **out);
void flatnet_destroy(FlatNet *net);
int flatnet_forward_backward(FlatNet *net, const float *x, const int64_t *labels, int n, int train, float sign,
                             float *loss_dev, float *logits_dev, cudaStream_t st);
int flatnet_forward(FlatNet *net, const float *x, int n, float *logits_dev, cudaStream_t st);
}  // namespace salun
