// psy_launch.h — kernel argument block shared by the host library (psy_capi.cu) and the generated model TU.
#pragma once

struct PsyInst {
  int person;  // person index (0-based position in the dataset, not the id)
  int code;    // >= 0: (theta_index << 1) | side, side 0 = theta-eps, 1 = theta+eps (parameter set 0);
               // < 0 : unperturbed instance of parameter set (-1 - code)  (-1 = the base instance of set 0)
};

struct PsyLaunch {
  const double* obs;            // [rows][M] row-major: one time point = M contiguous doubles; NaN = missing
  const double* dt;             // [rows]
  const int* ksteps;            // [rows] RK4 steps per row by odeint's integrate_const rule (host-computed, exact IEEE)
  const long long* row_off;     // [N+1] first row of each person
  const int* person_id;         // [N]   id value handed to the user equations as `person`
  const int* theta_index;       // [N][F] free slot -> theta index
  const double* theta;          // [n_sets][P]  (n_sets > 1 only for psy_fit_many)
  const PsyInst* inst;          // [n_inst]
  double* f_out;                // -2LL of each instance, at out_idx[i] (or i when out_idx is null)
  const int* out_idx;           // [n_inst] or null
  double* latent;               // [rows][N] row-major or null; written by base instances only
  double* pred;                 // [rows][M] row-major or null; written by base instances only
  double eps;                   // finite-difference step of this launch
  double h;                     // integrator step (rk4) / initial step (dopri5)
  double c, sqrtc, wm0, wm1, wc0, wc1;  // unscented-transform constants
  int n_inst;
  int skip_update;
  int n_theta;                  // P
};
