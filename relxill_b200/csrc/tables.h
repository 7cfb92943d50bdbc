// tables.h — FITS tables -> host arrays -> HBM-resident DevTables.
#pragma once
#include <string>
#include <vector>

#include "common.h"

namespace rx {

struct XillHost {
  bool loaded = false;
  int npar = 0, nvals[6] = {0}, pindex[6] = {0}, n_ener = 0, n_incl = 0, stride = 0;
  int xc_first = 0, xc_n = 0, xc_stride = 0;   // convolution bins overlapping the table grid (XillDev)
  bool has_conv_copy = false;                  // the fp64 convolution-grid copy of the rows is in HBM
  size_t bytes_c = 0;
  long nnodes = 0;
  std::vector<float> vals[6];
  std::vector<double> ener;
  size_t bytes = 0;
};

class Tables {
 public:
  ~Tables();
  // loads rel + (optionally) lp / rrad / xillver tables that exist in dir; returns "" or an error message
  std::string load(const std::string &dir);
  // lazily make sure the table needed by a model flavour is there
  std::string require(bool lp, bool rrad, int xtab);   // xtab: XT_* or XT_NONE
  std::string require_xill_only(int xtab);  // standalone xillver models need no relativistic table
  const DevTables &dev() const { return dt_; }
  const std::vector<double> &rr_spins() const { return rr_spin_; }
  const XillHost &xill_host(int xtab) const { return xh_[xtab]; }
  bool has_rel() const { return have_rel_; }
  const std::vector<double> &econv() const { return econv_; }
  size_t device_bytes() const { return dev_bytes_; }
  double conv_cf_deviation() const { return conv_cf_dev_; }
  // build the convolution-grid copy of the xillver tables that are loaded from now on (default on)
  void set_conv_grid_copy(bool on) { conv_grid_copy_ = on; }   // max |E_mid/dE / const - 1| on the convolution grid

 private:
  std::string dir_;
  DevTables dt_{};
  bool have_rel_ = false, have_lp_ = false, have_rr_ = false, have_fixed_ = false, have_nth_ = false;
  XillHost xh_[XT_COUNT];
  std::vector<double> rr_spin_, econv_, ecoarse_;
  std::vector<void *> allocs_;
  size_t dev_bytes_ = 0;
  double conv_cf_dev_ = 0.0;
  bool conv_grid_copy_ = true;
  bool upload_failed_ = false;   // a cudaMalloc / cudaMemcpy of a table array failed

  template <class T> const T *upload(const std::vector<T> &v);
  template <class T> const T *upload(const T *p, size_t n);
  void load_fixed();
  void load_nthcomp();
  std::string load_rel();
  std::string load_lp();
  std::string load_rrad();
  std::string load_xill(int which);
};

}  // namespace rx
