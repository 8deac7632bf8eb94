// dd.cuh — double-double arithmetic and the p1 probability, shared by host and device code.
//
// p1(k, d, n) = P(at most k-d distinct values in k draws with replacement from n)
// (enclone_help/src/help1.rs:292-298) is what join_one multiplies by N to get the score.
// The reference evaluates 1 - sum_{u=k-d+1..k} S(k,u) n!/(n-u)!/n^k in double-double (qd).
// Here the same sum is evaluated in closed form per n:
//     T_j(k) = S(k,k-j) * n^{\underline{k-j}} / n^k = B_j(k) * F(k-j),   j = k-u = 0..d-1
//     B_j(k) = S(k,k-j)/n^j        B_j(k) = ((k-j)/n) B_{j-1}(k-1) + B_j(k-1),  B_0 = 1
//     F(u)   = prod_{v<u} (1 - v/n)
//     p1     = 1 - T_0 - T_1 - ... - T_{d-1}          (subtracted in this order, in double-double)
// so a whole [d][k] table for one n costs O(K*D) double-double operations — cheap enough to
// build on the device — instead of O(d*k) per pair.  All operations use explicit fma and
// must be compiled without floating-point contraction (nvcc -fmad=false; gcc -ffp-contract=off)
// so that the host function ejoin_p1() and the device table hold identical doubles.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
#endif

struct dd_t { double hi, lo; };

DD_HD double dd_fma(double a, double b, double c) { return fma(a, b, c); }
DD_HD dd_t dd_make(double hi, double lo) { dd_t r; r.hi = hi; r.lo = lo; return r; }
DD_HD dd_t dd_from(double a) { return dd_make(a, 0.0); }
DD_HD dd_t dd_two_sum(double a, double b) {
    double s = a + b, bb = s - a;
    return dd_make(s, (a - (s - bb)) + (b - bb));
}
DD_HD dd_t dd_quick_two_sum(double a, double b) {
    double s = a + b;
    return dd_make(s, b - (s - a));
}
DD_HD dd_t dd_two_prod(double a, double b) {
    double p = a * b;
    return dd_make(p, dd_fma(a, b, -p));
}
DD_HD dd_t dd_add(dd_t a, dd_t b) {
    dd_t s = dd_two_sum(a.hi, b.hi), t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick_two_sum(s.hi, s.lo);
}
DD_HD dd_t dd_sub(dd_t a, dd_t b) { return dd_add(a, dd_make(-b.hi, -b.lo)); }
DD_HD dd_t dd_mul(dd_t a, dd_t b) {
    dd_t p = dd_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return dd_quick_two_sum(p.hi, p.lo);
}
DD_HD dd_t dd_mul_d(dd_t a, double b) {
    dd_t p = dd_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return dd_quick_two_sum(p.hi, p.lo);
}
DD_HD dd_t dd_div(dd_t a, dd_t b) {
    double q1 = a.hi / b.hi;
    dd_t r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    dd_t q = dd_quick_two_sum(q1, q2);
    return dd_add(q, dd_from(q3));
}
DD_HD double dd_to_double(dd_t a) { return a.hi + a.lo; }

// Table geometry: P1[d][k] for d < P1_DCAP, k < P1_KCAP, one table per n.
#define P1_DCAP 160
#define P1_KCAP 1024

// One step of the B recurrence: B_j(k) from row k-1.  prev_jm1 = B_{j-1}(k-1), prev_j = B_j(k-1).
DD_HD dd_t p1_b_step(int k, int j, dd_t prev_jm1, dd_t prev_j, double n) {
    if (j == 0) return dd_from(1.0);
    if (j > k) return dd_from(0.0);
    dd_t coef = dd_div(dd_from((double)(k - j)), dd_from(n));
    return dd_add(dd_mul(coef, prev_jm1), prev_j);
}
// F(u+1) = F(u) * (1 - u/n) = F(u) * (n-u)/n
DD_HD dd_t p1_f_step(dd_t f, int u, double n) {
    return dd_mul(f, dd_div(dd_from(n - (double)u), dd_from(n)));
}
