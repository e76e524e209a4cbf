// svgp.cu -- K7/K8 (forward + hand-derived backward) -- entry points.
#include "svgp.cuh"

extern "C" {
int mfgp_svgp_elbo_grad(mfgp_handle* h, const mfgp_svgp_cfg*, const double*, const double*, const double*,
                        const double*, const double*, const double*, const double*, double, double*, double*,
                        double*, double*, double*, double*, double*, double*) {
    return mfgp_fail(h, MFGP_ERR_UNSUPPORTED, "mfgp_svgp_elbo_grad: not built yet");
}
int mfgp_svgp_predict(mfgp_handle* h, const mfgp_svgp_cfg*, const double*, int, const double*, const double*,
                      const double*, const double*, const double*, double*, double*) {
    return mfgp_fail(h, MFGP_ERR_UNSUPPORTED, "mfgp_svgp_predict: not built yet");
}
}
