/*
 * perturb_benchtools.h -- measurement and test hooks (libperturb_benchtools.so). NOT part of the
 * drop-in ABI and not linked into libperturb_gpu.so: the benchmark's reference arm, which must
 * not map the product library, takes its synthetic catalogues (perturb_synth.h, same library)
 * from here, and the test-suite probes the device math routines through it.
 * Status codes are those of perturb_gpu.h (0 ok, negative failure).
 */
#ifndef PERTURB_BENCHTOOLS_H
#define PERTURB_BENCHTOOLS_H

#include <stddef.h>

#include "perturb_synth.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FP64 roofline denominator: DFMA-only microbenchmark on `device` (no memory traffic),
 * best of 5 after a warm-up; *tflops counts a DFMA as 2 flops. Used by bench.py. */
int perturb_gpu_measure_fp64_peak(int device, double *tflops);

/* Test hook: out[i] = f(x[i], y[i]) for the device math routine number `fn` (see
 * perturb_b200/csrc/math_probe.cu); host pointers. Lets the test-suite bound the error of
 * the FP64 fast paths against correctly rounded host values. */
int perturb_gpu_math_probe(int fn, const double *x, const double *y, double *out, size_t n);

/* Raw device -> pinned host copy rate of `device`: `reps` timed copies of `bytes` bytes, each
 * one cudaMemcpyAsync (after an untimed one); *gbs_best = the fastest copy, *gbs_mean (may be
 * NULL) = all bytes / all time, in GB/s. The denominator of the end-to-end (host buffer)
 * figures: 52 bytes per evaluation have to cross this link. Ranks that call it at the same
 * time measure the rate they get while the others copy too. */
int perturb_gpu_measure_d2h(int device, size_t bytes, int reps, double *gbs_best,
                            double *gbs_mean);

/* The same from ONE process over devices 0 .. n_devices-1 at once, pinned buffers allocated with
 * `host_flags` (cudaHostAlloc flags: 0 default, 1 portable, 4 write-combined); *gbs_total = all
 * bytes / wall time with every copy in flight together. */
int perturb_gpu_measure_d2h_multi(int n_devices, size_t bytes, int reps, unsigned host_flags,
                                  double *gbs_total);

#ifdef __cplusplus
}
#endif
#endif /* PERTURB_BENCHTOOLS_H */
