// Shear-function hooks: scalar host functions rate(t), strain(t), offset, and the wrapped-strain
// variant that drives the box tilt.  C mirrors of the reference's C++ classes
// (PSEv1/ShearFunction.h:19-36, PSEv1/SpecificShearFunction.h:16-223,
// PSEv1/VariantShearFunction.h:46-48, PSEv1/VariantShearFunction.cc:17-43).
// Kept quirks: pi literal 3.1415926536, single-precision logf inside the chirp, and
// (timestep - offset) evaluated in unsigned arithmetic.
#include <math.h>
#include <stdint.h>
#include "../../include/pse_b200.h"

struct pse_shear {
    int kind;
    double a[4];
    uint32_t offset;
    double dt;
    pse_shear *base, *window;
};

static const double kPi = 3.1415926536;

extern "C" pse_shear* pse_shear_create(int kind, const double* args, int nargs, uint32_t offset, double dt) {
    static const int need[] = {0, 1, 2, 4, 2};
    if (kind < 0 || kind > 4 || nargs != need[kind]) return nullptr;
    pse_shear* s = new pse_shear();
    s->kind = kind; s->offset = offset; s->dt = dt; s->base = s->window = nullptr;
    for (int i = 0; i < 4; ++i) s->a[i] = i < nargs ? args[i] : 0.0;
    return s;
}
extern "C" pse_shear* pse_shear_create_windowed(pse_shear* base, pse_shear* window) {
    if (!base || !window) return nullptr;
    pse_shear* s = new pse_shear();
    s->kind = 5; s->offset = 0; s->dt = 0; s->base = base; s->window = window;
    for (int i = 0; i < 4; ++i) s->a[i] = 0;
    return s;
}
extern "C" void pse_shear_destroy(pse_shear* s) { delete s; }

static double elapsed(const pse_shear* s, uint32_t t) { return (double)(uint32_t)(t - s->offset) * s->dt; }  // unsigned wrap kept

// Tukey window: relative time in [0,1], cosine lobes of relative width param/2 at both ends
static double tukey_rel(const pse_shear* s, uint32_t t) { return (uint32_t)(t - s->offset) * s->dt / s->a[0]; }

extern "C" double pse_shear_rate(const pse_shear* s, uint32_t t) {
    if (!s) return 0.0;
    switch (s->kind) {
        case 1: return s->a[0];                                                    // steady: rate
        case 2: return s->a[0] * cos(s->a[1] * 2 * kPi * elapsed(s, t));           // sine: max_rate, frequency
        case 3: {                                                                  // chirp: amp, w0, wf, periodT
            const double amp = s->a[0], w0 = s->a[1], wf = s->a[2], T = s->a[3];
            const double lg = logf((float)(wf / w0));
            const double ex = exp(s->dt * (uint32_t)(t - s->offset) * lg / T);
            const double omega = w0 * ex, phase = T * w0 / lg * (ex - 1);
            return amp * omega * cos(phase);
        }
        case 4: {                                                                  // tukey: periodT, param
            const double T = s->a[0], p = s->a[1], w = 2 * kPi / p, rel = tukey_rel(s, t);
            if (rel <= 0 || rel >= 1) return 0;
            if (rel >= p / 2 && rel <= 1 - p / 2) return 0;
            if (rel < 0.5) return -(sin(w * (rel - p / 2))) / 2 * w / T;
            return -(sin(w * (rel - 1 + p / 2))) / 2 * w / T;
        }
        case 5:                                                                    // product rule
            return pse_shear_rate(s->base, t) * pse_shear_strain(s->window, t) +
                   pse_shear_strain(s->base, t) * pse_shear_rate(s->window, t);
        default: return 0.0;
    }
}

extern "C" double pse_shear_strain(const pse_shear* s, uint32_t t) {
    if (!s) return 0.0;
    switch (s->kind) {
        case 1: return s->a[0] * (uint32_t)(t - s->offset) * s->dt;
        case 2: return s->a[0] * sin(s->a[1] * 2 * kPi * elapsed(s, t)) / s->a[1] / 2 / kPi;
        case 3: {
            const double amp = s->a[0], w0 = s->a[1], wf = s->a[2], T = s->a[3];
            const double lg = logf((float)(wf / w0));
            const double phase = T * w0 / lg * (exp(s->dt * (uint32_t)(t - s->offset) * lg / T) - 1);
            return amp * sin(phase);
        }
        case 4: {
            const double p = s->a[1], w = 2 * kPi / p, rel = tukey_rel(s, t);
            if (rel <= 0 || rel >= 1) return 0;
            if (rel >= p / 2 && rel <= 1 - p / 2) return 1;
            if (rel < 0.5) return (1 + cos(w * (rel - p / 2))) / 2;
            return (1 + cos(w * (rel - 1 + p / 2))) / 2;
        }
        case 5: return pse_shear_strain(s->base, t) * pse_shear_strain(s->window, t);
        default: return 0.0;
    }
}

extern "C" uint32_t pse_shear_offset(const pse_shear* s) {
    if (!s) return 0;
    return s->kind == 5 ? pse_shear_offset(s->base) : s->offset;
}

extern "C" double pse_shear_variant_value(const pse_shear* s, uint32_t total, double vmin, double vmax, uint32_t t) {
    const uint32_t off = pse_shear_offset(s);
    const double range = vmax - vmin;
    auto wrap = [&](double v) { return v - range * floor((v - vmin) / range); };
    if (t < off) return 0;
    if (t >= off + total) return wrap(pse_shear_strain(s, off + total));
    return wrap(pse_shear_strain(s, t));
}
