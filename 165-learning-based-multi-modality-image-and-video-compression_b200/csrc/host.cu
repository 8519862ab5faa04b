// host.cu -- host-side pieces of the C ABI: version, thread-local error string, launch counter and
// pmf_to_quantized_cdf (compressai/cpp_exts/ops/ops.cpp:40-109; once-per-model work in update()).
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mmc {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

}  // namespace mmc

extern "C" {

int mmc_version(void) { return MMC_VERSION; }
const char *mmc_last_error(void) { return mmc::g_err; }
int64_t mmc_launch_count(void) { return mmc::g_launches; }
void mmc_reset_launch_count(void) { mmc::g_launches = 0; }

int mmc_pmf_to_quantized_cdf_host(const float *pmf, int n, int precision, uint32_t *cdf)
{
    MMC_CHECK_ARG(pmf && cdf && n >= 1 && precision >= 1 && precision <= 30, "mmc_pmf_to_quantized_cdf_host: bad argument");
    for (int i = 0; i < n; ++i) {
        if (pmf[i] < 0.0f || !std::isfinite(pmf[i])) {
            mmc::set_error("Invalid `pmf`, non-finite or negative element found: %f", pmf[i]);
            return MMC_EDOMAIN;
        }
    }
    // 1. scale to integer frequencies, 2. renormalise to 2^precision, 3. prefix sum
    const float scale = (float)(1 << precision);
    cdf[0] = 0;
    uint32_t total = 0;
    for (int i = 0; i < n; ++i) {
        cdf[i + 1] = (uint32_t)std::round(pmf[i] * scale);
        total += cdf[i + 1];
    }
    if (total == 0) {
        mmc::set_error("Invalid `pmf`: at least one element must have a non-zero probability.");
        return MMC_EDOMAIN;
    }
    uint32_t run = 0;
    for (int i = 0; i <= n; ++i) {
        run += (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
        cdf[i] = run;
    }
    cdf[n] = 1u << precision;
    // 4. every symbol needs a non-zero frequency: take one count from the cheapest donor
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t best = ~0u;
        int donor = -1;
        for (int j = 0; j < n; ++j) {
            uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < best) { best = f; donor = j; }
        }
        if (donor < 0) {
            mmc::set_error("pmf_to_quantized_cdf: no symbol can donate frequency");
            return MMC_EDOMAIN;
        }
        if (donor < i) for (int j = donor + 1; j <= i; ++j) cdf[j]--;
        else           for (int j = i + 1; j <= donor; ++j) cdf[j]++;
    }
    return MMC_OK;
}

}  // extern "C"
