// conv_tc.cu -- tensor-core implicit-GEMM convolution (placeholder until the tcgen05 kernel lands).
#include "common.cuh"

extern "C" {

int mmc_conv_pack_weights(const mmc_conv_desc *d, const float *w, void *w_packed, size_t *bytes, void *stream)
{
    (void)d; (void)w; (void)w_packed; (void)bytes; (void)stream;
    mmc::set_error("mmc_conv_pack_weights: tensor-core path not built yet");
    return MMC_EUNSUPPORTED;
}

int mmc_conv_forward_tc(const mmc_conv_desc *d, const void *x, const void *w_packed, const float *bias,
                        const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream)
{
    (void)d; (void)x; (void)w_packed; (void)bias; (void)beta_eff; (void)gamma_eff_bf16; (void)y; (void)y2; (void)stream;
    mmc::set_error("mmc_conv_forward_tc: tensor-core path not built yet");
    return MMC_EUNSUPPORTED;
}

}
