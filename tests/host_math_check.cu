// Host build of ml_unigen_b200/csrc/ug_math.cuh: the kernel's arithmetic source, executed on
// the CPU so tests can compare it bit for bit with oracle/decode_step_oracle.c without a GPU.
#include "ug_math.cuh"
extern "C" {
float ugk_host_expf_exact(float d) { return ug::expf_exact(d); }
float ugk_host_expf_fast(float d) { return ug::expf_fast(d); }
float ugk_host_logf(float x) { return ug::logf_det(x); }
float ugk_host_race_q(unsigned long long seed, unsigned long long offset, unsigned long long row, unsigned i) { return ug::race_q(i, row, seed, offset); }
float ugk_host_select_u(unsigned long long seed, unsigned long long offset, unsigned long long b, unsigned n) { return ug::select_u(n, b, seed, offset); }
float ugk_host_confidence(float p, float u, float t) { return ug::confidence(p, u, t); }
unsigned ugk_host_cfg_mix(unsigned c, unsigned u, float s) { return ug::cfg_mix_bits(c, u, s); }
unsigned ugk_host_ordered_key(float f) { return ug::ordered_key(f); }
void ugk_host_expf_scan(unsigned lo_bits, unsigned hi_bits, unsigned stride, float (*oracle)(float), unsigned long long* mism_exact, unsigned long long* mism_fast) {
    unsigned long long me = 0, mf = 0;
    for (unsigned long long u = lo_bits; u <= hi_bits; u += stride) {
        float d = ug_host_u2f((unsigned)u);
        float o = oracle(d);
        if (ug_host_f2u(ug::expf_exact(d)) != ug_host_f2u(o)) ++me;
        if (d >= -86.0f && ug_host_f2u(ug::expf_fast(d)) != ug_host_f2u(o)) ++mf;
    }
    *mism_exact = me; *mism_fast = mf;
}
void ugk_host_logf_scan(unsigned lo_bits, unsigned hi_bits, unsigned stride, float (*oracle)(float), unsigned long long* mism) {
    unsigned long long m = 0;
    for (unsigned long long u = lo_bits; u <= hi_bits; u += stride) {
        float x = ug_host_u2f((unsigned)u);
        if (ug_host_f2u(ug::logf_det(x)) != ug_host_f2u(oracle(x))) ++m;
    }
    *mism = m;
}
}
