// backbone_ops.h -- placeholder for the fused row-wise backbone kernels (RMSNorm, RoPE, SwiGLU)
#pragma once
