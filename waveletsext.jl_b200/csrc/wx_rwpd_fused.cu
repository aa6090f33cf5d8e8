// wx_rwpd_fused.cu -- fused 1-D swpd / acwpd kernel (placeholder until the fused kernel lands: reports "not handled"
// so that wx_rwt.cu runs the per-depth path).
#include "wx_steps.cuh"

template <typename T>
int wx_rwpd1d_fused(int ac, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    return WX_OK;
}
template int wx_rwpd1d_fused<double>(int, double *, const double *, long, int, long, const Taps<double> &, cudaStream_t, bool *);
template int wx_rwpd1d_fused<float>(int, float *, const float *, long, int, long, const Taps<float> &, cudaStream_t, bool *);
