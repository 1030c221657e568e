// kt_kernels.h -- host-visible launcher table.  kt_kernels.cu is compiled once per arithmetic mode and
// exports one table per mode; kt_api.cu picks the table from kt_run::math_mode.
#pragma once
#include <cuda_runtime.h>

#include "kt_types.h"

namespace kt {

struct LaunchCfg {
    cudaStream_t stream;
    int sm_count;
    int min_blocks;   // CTAs of 128 threads per SM the trace kernel is compiled for (3, 4 or 5); 0 = default
    int carveout;     // preferred shared-memory carve-out in percent, -1 = leave it to the driver
};

// device pointers of the nine caustic maps (+ the snapshot of sign_J the suppression pass reads, + 7 counters:
// disc_count, horizon, rlim, steplim, outofrange, other, suppressed)
struct CausticMaps {
    double *det_J, *sign_J, *radius, *phi, *x_disc, *y_disc, *redshift, *sign_snap;
    int *order, *hit;
    unsigned long long* cnt;
};

struct KernelTable {
    cudaError_t (*trace_rayset)(const LaunchCfg&, const TraceParams&, kt_ray* rays, double* counters);
    cudaError_t (*trace_emis)(const LaunchCfg&, const TraceParams&, const PointSourceFrame&, const kt_emis_bins&, double log_dr,
                              double* hist);
    cudaError_t (*trace_image)(const LaunchCfg&, const TraceParams&, const PlaneFrame&, const kt_image_bins&, int do_range_phi,
                               double* maps);
    cudaError_t (*trace_planerays)(const LaunchCfg&, const TraceParams&, const PlaneFrame&, const kt_image_bins&,
                                   int do_range_phi, kt_ray_min* out, double* counters);
    cudaError_t (*rayset_reset)(const LaunchCfg&, kt_ray*, long long n);
    cudaError_t (*init_pointsource)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&, const PointSourceFrame&);
    cudaError_t (*init_imageplane)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&, const PlaneFrame&);
    cudaError_t (*redshift_start)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&, double V, int reverse, int projradius);
    cudaError_t (*redshift_V)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&, double V, int reverse, int projradius,
                              int motion);
    cudaError_t (*redshift_dest)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&, int reverse);
    cudaError_t (*range_phi)(const LaunchCfg&, kt_ray*, long long n, double lo, double hi);
    cudaError_t (*calc_momentum)(const LaunchCfg&, kt_ray*, long long n, const TraceParams&);
    cudaError_t (*bin_emissivity)(const LaunchCfg&, const kt_ray*, long long n, const kt_emis_bins&, double log_dr, double* hist);
    cudaError_t (*bin_image)(const LaunchCfg&, const kt_ray*, long long n, const kt_image_bins&, double* maps);
    cudaError_t (*dfma_peak)(const LaunchCfg&, double* out, int blocks, int threads, int iters);
    // caustic_discplane.cpp:219-334 / 351-448 (pass 0: centre rays, pass 1: Jacobian, pass 2: suppression)
    cudaError_t (*caustic_pass)(const LaunchCfg&, int pass, const kt_ray_min* rays, long long n_live, int plane_Ny, int bundles,
                                double eps_x, double eps_y, const kt_caustic_cfg&, const CausticMaps&);
    cudaError_t (*trace_paths)(const LaunchCfg&, const TraceParams&, kt_ray* rays, double* counters, const kt_path_opts&,
                               long long* counts, const long long* offsets, double* points);
};

}  // namespace kt

extern "C" const kt::KernelTable* kt_kernel_table_exact();
extern "C" const kt::KernelTable* kt_kernel_table_fast();
