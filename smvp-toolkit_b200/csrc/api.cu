// api.cu -- the small C-ABI entry points that are not kernels: error text, device probe,
// launch counter and the iteration statistics of struct _time_data_ (main-cli.c:87-95, :428-456).
#include "common.cuh"

#include <math.h>

using namespace smvp;

extern "C" const char *smvp_strerror(int code)
{
    switch (code)
    {
    case SMVP_OK:
        return "ok";
    case SMVP_E_ARG:
        return "invalid argument";
    case SMVP_E_ALLOC:
        return "allocation failed";
    case SMVP_E_CUDA:
        return "CUDA error (no device, or a runtime/launch failure; see smvp_last_cuda_error)";
    case SMVP_E_RANGE:
        return "matrix coordinate out of range";
    case SMVP_E_TOOBIG:
        return "problem too large for int32 offsets";
    default:
        return "unknown error";
    }
}

extern "C" const char *smvp_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int smvp_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return n;
}

extern "C" int64_t smvp_launch_count(void) { return (int64_t)g_launches.load(); }

// total / mean / min / max and the POPULATION standard deviation the reference intends
// (calcStDevDouble, main-cli.c:114-130: sqrt(sum((t - mean)^2) / n); its accumulators are
// uninitialised there, U11 -- zero is the intended start).
extern "C" int smvp_time_stats(const double *ms_each, int n, smvp_time_stats_t *out)
{
    if (!out || n < 0 || (n > 0 && !ms_each))
        return SMVP_E_ARG;
    double total = 0.0, mn = 0.0, mx = 0.0;
    for (int i = 0; i < n; i++)
    {
        total += ms_each[i];
        if (i == 0 || ms_each[i] < mn)
            mn = ms_each[i];
        if (i == 0 || ms_each[i] > mx)
            mx = ms_each[i];
    }
    const double mean = n > 0 ? total / n : 0.0;
    double ss = 0.0;
    for (int i = 0; i < n; i++)
        ss += (ms_each[i] - mean) * (ms_each[i] - mean);
    out->time_total = total;
    out->time_avg = mean;
    out->time_stdev = n > 0 ? sqrt(ss / n) : 0.0;
    out->time_min = mn;
    out->time_max = mx;
    return SMVP_OK;
}
