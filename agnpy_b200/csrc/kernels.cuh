// Host-side launchers of the kernel families (definitions in synch_ssc.cu, ec.cu, tau.cu).
#pragma once
#include "common.cuh"

namespace agb {

void launch_prep_gamma(const agnpy_model_t* models, int n_models, int ne_kind, const double* gamma,
                       int n_gamma, const double* ne_table, const double* ssa_table, int mode,
                       double* tab, cudaStream_t stream);

int launch_synch(const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
                 const double* tab, int n_gamma, int ssa, int integ, int out_mode, double* out,
                 double* tau_out, double* eps_out, cudaStream_t stream, int nu_stride = 0,
                 int proton = 0);

int launch_ssc_loss_rate(int n_models, const double* gamma_eval, int n_eval, const double* U_seed,
                         const double* eps_seed, int n_seed, double* out, cudaStream_t stream);

// `scratch`: ssc_scratch_doubles(n_gamma, n_seed) doubles per model, 32-byte aligned
size_t ssc_scratch_doubles(int n_gamma, int n_seed);
int launch_ssc(const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
               const double* tab, int n_gamma, const double* U_seed, const double* eps_seed,
               int n_seed, int integ, double* out, double* scratch, cudaStream_t stream);

// bytes of scratch the EC / tau pipelines need for `n_models` models
size_t ec_workspace_bytes(int variant, int n_models, int n_nu, int n_gamma, int n_mu, int n_phi, int integ);
int run_ec(int variant, const agnpy_model_t* models, int n_models, int ne_kind, const double* nu,
           int n_nu, const double* gamma, int n_gamma, const double* mu, int n_mu,
           const double* phi, int n_phi, const double* ne_table, int integ, double* sed,
           double* ws, cudaStream_t stream);

size_t tau_workspace_bytes(int variant, int n_models, int n_nu, int n_a, int n_phi, int n_l);
int run_tau(int variant, const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
            const double* mu, int n_a, const double* phi, int n_phi, const double* l_unit, int n_l,
            double* tau, double* ws, cudaStream_t stream);

// opacity on the blob's own synchrotron photons (targets_absorption.py:629-694)
size_t tau_synch_workspace_bytes(int n_models, int n_nu, int n_gamma, int n_s);
int run_tau_synch(const agnpy_model_t* models, int n_models, int ne_kind, const double* nu, int n_nu,
                  const double* gamma, int n_gamma, const double* ne_table, const double* ssa_table,
                  int n_s, double margin_low, int integ, double* tau, double* ws, cudaStream_t stream);

int run_trapz_loglog(const double* y, const double* x, long long n_rows, int n_x, int intervals,
                     double* out, cudaStream_t stream);

int run_bb(int variant, const agnpy_model_t* models, int n_models, const double* nu, int n_nu,
           const double* nu_norm, int n_norm, int n_R, double* sed, cudaStream_t stream);

int run_target_u(int variant, const agnpy_model_t* models, int n_models, const double* Gamma,
                 const double* r, int n_r, int n_mu, double* u, cudaStream_t stream);
int run_ebl(int n_models, const double* z, const double* nu, int n_nu, const double* e_ref, int n_e,
            const double* z_ref, int n_z, const double* values, int multiply, double* out,
            cudaStream_t stream);

}  // namespace agb
