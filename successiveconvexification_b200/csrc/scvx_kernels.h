// scvx_kernels.h — host-side launchers of the CUDA kernels (internal to the shared library).
#pragma once
#include <cuda_runtime.h>
#include "scvx_common.cuh"

cudaError_t scvx_launch_dualwarp(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s);
cudaError_t scvx_launch_fin_force(const ScvxTables& lift_tb, const ScvxTables& drag_tb, const double* mach, const double* defl,
                                  int n, double* out_lift, double* out_drag, cudaStream_t s);
cudaError_t scvx_launch_dualwarp_fins(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s);
cudaError_t scvx_launch_predict(const ScvxBatch& bt, const ScvxTables& tb, cudaStream_t s);
cudaError_t scvx_launch_prefilter(const double* d_samples, int n1, int n2, double* d_tmp, double* d_coef,
                                  const double* d_cp, cudaStream_t s);
cudaError_t scvx_launch_defect_cost(const double* X, const double* lin_err, int n_nodes, int B, double wNu,
                                    double* out_defect, double* out_cost, cudaStream_t s);
cudaError_t scvx_launch_linear_points(const double* rIi, const double* vIi, const double* mwet, double mwet_shared, double mdry,
                                      const double* rIf, const double* vIf, double g, int K, int B, double* X, double* U,
                                      cudaStream_t s);
cudaError_t scvx_launch_socp_values(const double* blocks, const double* lin_err, const double* tlb, int n_nodes, int B,
                                    double* vals, double* rhs, cudaStream_t s);
cudaError_t scvx_launch_dispersed_setup(const scvx_dim_problem& base, const double* rIi, const double* vIi, const double* mwet,
                                        int B, double* X, double* U, double* sigma, double* scales, scvx_probinfo* P0,
                                        scvx_probinfo* P1, cudaStream_t s);
cudaError_t scvx_launch_compact_pack(const double* blocks, long n_intervals, int record_doubles, double* out, int sm_count,
                                     cudaStream_t s);

// STAGED path (scvx_kernels_staged.cu): value kernel + persistent tangent kernel, chunked over a scratch buffer.
cudaError_t scvx_staged_init();          // kernel attributes of the current device (once per context and device)
size_t scvx_staged_scratch_bytes(int npts, int chunk_intervals);
int scvx_staged_chunk_intervals(int sm_count);
// shared_params != nullptr: the record travels in the kernel arguments (constant bank).  sweep == false: every trajectory
// uses it as is; sweep == true: the records of bt.P differ from it in `a` and `Tmin` only, which the kernels read per
// trajectory from bt.P (mass / thrust-bound sweeps, BASELINE.json configs[3]).
cudaError_t scvx_launch_staged(const ScvxBatch& bt, const ScvxTables& tb, bool any_aero, const scvx_probinfo* shared_params,
                               bool sweep, void* scratch, int chunk_intervals, int sm_count, cudaStream_t s, int* launches);

