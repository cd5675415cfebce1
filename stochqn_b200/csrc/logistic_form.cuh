// logistic_form.cuh - the two binary-logistic conventions and the per-row weight they share; used by the bundled
// callbacks (callbacks.cu) and by the fused mini-batch kernel of the device-side request loop (kernels_fit.cuh).
#pragma once

enum { LG_GRAD = 0, LG_HVP = 1, LG_LOSS = 2 };

// Two conventions share the kernels:
//   sk = 0  R/logistic.R:1-37: y in {0,1}, weighted MEANS, penalty lambda*|w|^2 on every coefficient (the intercept is a
//           column of X);
//   sk = 1  scikit-learn (<= 1.0) _logistic_loss_and_grad / _logistic_grad_hess, which the reference's Python layer calls
//           (stochqn/_logistic.py:23-30): y in {-1,+1}, weighted SUMS, penalty alpha/2*|w[:ncols]|^2, and with icpt = 1
//           an unpenalised intercept stored LAST in w (w has ncols + 1 entries; z = x'w[:ncols] + w[ncols]).
struct LgForm { int sk; int icpt; };

__device__ __forceinline__ double lg_row_weight(int kind, const LgForm f, double z, double t, double yy, double wt)
{
    if (!f.sk) {
        const double p = 1.0 / (1.0 + exp(-z));
        if (kind == LG_GRAD) return (p - yy) * wt;
        if (kind == LG_HVP) return p * (1.0 - p) * wt * t;
        return -(yy * log(p) + (1.0 - yy) * log(1.0 - p)) * wt;
    }
    const double yz = yy * z;
    const double q = 1.0 / (1.0 + exp(-yz));
    if (kind == LG_GRAD) return wt * (q - 1.0) * yy;
    if (kind == LG_HVP) return wt * q * (1.0 - q) * t;
    return wt * (yz > 0 ? log1p(exp(-yz)) : -yz + log1p(exp(yz)));        // -log sigmoid(yz)
}
