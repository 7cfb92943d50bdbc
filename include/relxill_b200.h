/* relxill_b200.h — C ABI of librelxill_b200.so, the B200 (sm_100a) implementation of relxill's
 * spectrum-evaluation hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * Two groups of entry points:
 *
 *  (1) the XSPEC/ISIS local-model functions, drop-in for the symbols the reference's generated
 *      wrapper exports (reference src/create_wrapper_xspec.py:153-162, one per block of
 *      src/modelfiles/lmodel_relxill_public.dat, dispatched by xspec_C_wrapper_eval_model,
 *      src/LocalModel.cpp:143-160).  Same signature, same parameter order, same units.
 *
 *  (2) the batched entry points (new; north_star): N parameter vectors on one shared energy grid.
 *
 * There is no CPU fallback: every entry point runs the CUDA kernels and fails (non-zero return,
 * message on stderr, flux zeroed) if no device / tables are available.
 */
#ifndef RELXILL_B200_H_
#define RELXILL_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- (1) XSPEC local models
 * energy[Nflux+1] ascending bin edges (keV); parameter[] in lmodel.dat order including the
 * $switch entries, excluding norm; flux[Nflux] out (photons/cm^2/s/bin).  For the convolution
 * models ("con") flux is input and output.  spectrum, fluxError and init are ignored, as in the
 * reference.  Tables are read from $RELXILL_TABLE_PATH (or ./) on first use, like the reference
 * (src/relutility.c:320-328). */
#define RELXILL_B200_LMOD(name)                                                                            \
  void name(const double *energy, int Nflux, const double *parameter, int spectrum, double *flux,         \
            double *fluxError, const char *init)

RELXILL_B200_LMOD(lmodrelline);              /* lmodel_relxill_public.dat:1   relline      (10 par) */
RELXILL_B200_LMOD(lmodrelconv);              /* :13  relconv      (8)  */
RELXILL_B200_LMOD(lmodrellinelp);            /* :23  relline_lp   (10) */
RELXILL_B200_LMOD(lmodrelconvlp);            /* :35  relconv_lp   (9)  */
RELXILL_B200_LMOD(lmodrelxill);              /* :46  relxill      (13) */
RELXILL_B200_LMOD(lmodrelxilllp);            /* :61  relxilllp    (14) */
RELXILL_B200_LMOD(lmodrelxilldensnthcomp);   /* :96  relxillCp    (14) */
RELXILL_B200_LMOD(lmodrelxilllpdensnthcomp); /* :112 relxilllpCp  (17) */
RELXILL_B200_LMOD(lmodxillver);              /* :77  xillver      (7)  */
RELXILL_B200_LMOD(lmodxillverdensnthcomp);   /* :86  xillverCp    (8)  */
RELXILL_B200_LMOD(lmodxillverns);            /* :131 xillverNS    (7)  blackbody-irradiated table xillverNS-2.fits */
RELXILL_B200_LMOD(lmodrelxillns);            /* :140 relxillNS    (13) */
RELXILL_B200_LMOD(lmodxillverco);            /* lmodel_relxill_devel.dat:1  xillverCO (8)  table xillverCO.fits */
RELXILL_B200_LMOD(lmodrelxillco);            /* lmodel_relxill_devel.dat:11 relxillCO (14) */

/* ---------------------------------------------------------------- library state */
/* Select the CUDA device (-1: the current one) and load the tables from `table_dir` (NULL: $RELXILL_TABLE_PATH or
 * "./").  Optional: the first evaluation does it lazily on the current device — or, if the environment variable
 * RELXILL_B200_DEVICES is set ("all" or a count), on that many devices.  Returns 0 on success. */
int relxill_b200_init(const char *table_dir, int device);
/* Several devices in ONE process (north_star: "the parameter batch is sharded across the 8 GPUs of one box"): one
 * engine — tables replicated, own scratch arena and streams — on each of the devices 0 .. n_devices-1 (n_devices < 1:
 * all visible devices).  relxill_batch_eval and the lmod* symbols then shard every batch of >= 2 vectors per device
 * over the engines, one host thread per device, and every device copies its rows straight into the caller's flux
 * array: no collective is needed for host-buffer calls.  Prepared batches live on one engine
 * (relxill_b200_prepare_on). */
int relxill_b200_init_devices(const char *table_dir, int n_devices);
int relxill_b200_num_devices(void);
/* How relxill_batch_eval splits a batch over several devices: 0 (default) contiguous blocks, 1 round-robin rows — for
 * structured batches (parameter-grid sweeps) whose cost varies systematically along the batch.  Also settable with the
 * environment variable RELXILL_B200_INTERLEAVE=1. */
void relxill_b200_set_sharding(int interleave);
/* Free all device memory (tables, scratch) on every device. */
void relxill_b200_shutdown(void);
/* Programmatic override of the reference's RELXILL_NUM_RZONES environment variable (src/relutility.c:506-544):
 * n > 0 takes precedence over the variable, 0 removes the override.  Without an override the variable is read on
 * EVERY call, as are RELXILL_RETURNRAD_SWITCH, RELLINE_PHYSICAL_NORM, RELXILL_CONSTANT_DENSITY and
 * RELXILL_RENORMALIZE — the reference re-reads all of them per evaluation (src/relutility.c:372-396,506-544;
 * src/ModelDefinition.cpp:123-149; src/Relxill.cpp:241-278), and callers flip them mid-session. */
void relxill_b200_set_num_zones(int n);
/* Number of parameters of a model ("relxilllp", ...; XSPEC names), -1 if unknown. */
int relxill_b200_num_params(const char *model);
/* Default parameter vector (lmodel.dat column 3); returns the count or -1. */
int relxill_b200_default_params(const char *model, double *out);
/* Last error message of this thread ("" if none). */
const char *relxill_b200_last_error(void);

/* ---------------------------------------------------------------- (2) batched evaluation
 * model   XSPEC model name.
 * energy  [n_flux+1] shared bin edges (keV), host memory.
 * params  [n_vec][npar] row-major, host memory.
 * flux    [n_vec][n_flux] row-major, host memory (input as well for convolution models).
 * status  [n_vec] or NULL: 0 = ok, >0 = this vector's parameters were rejected / evaluation
 *         failed (its flux row is zeroed), mirroring the reference's per-call failure.
 * Returns 0 if the batch ran (individual vectors may still carry a status), <0 on a library
 * error (no device, tables missing, unknown model). */
int relxill_batch_eval(const char *model, const double *energy, int n_flux, const double *params, long n_vec,
                       double *flux, int *status);

/* Same, but the result stays on the device: d_flux is a device pointer to [n_vec][n_flux]
 * doubles (e.g. a torch tensor's data_ptr) and the work is enqueued on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream).  Used for the multi-GPU gather. */
int relxill_batch_eval_device(const char *model, const double *energy, int n_flux, const double *params,
                              long n_vec, double *d_flux, int *status, void *stream);

/* Prepared batches: interpret + upload once, evaluate many times with everything resident in
 * HBM (what bench.py times as `value`). */
typedef struct relxill_b200_batch relxill_b200_batch;
relxill_b200_batch *relxill_b200_prepare(const char *model, const double *energy, int n_flux,
                                         const double *params, long n_vec);
/* ... on the engine with index `device_index` (0 .. relxill_b200_num_devices()-1); relxill_b200_prepare uses engine 0 */
relxill_b200_batch *relxill_b200_prepare_on(int device_index, const char *model, const double *energy, int n_flux,
                                            const double *params, long n_vec);
/* Enqueues the kernels of the batch on `stream` (a cudaStream_t of the batch's device) and returns; d_flux is device
 * memory of that device.  Asynchronous with respect to the host: synchronise the stream (or call
 * relxill_b200_batch_status, which waits) before reading the result.  Runs that share an engine are ordered one after
 * the other on the device whatever streams they are given (they share the engine's scratch arena). */
int relxill_b200_run(relxill_b200_batch *b, double *d_flux, void *stream);
/* Device-resident state cache (the reference's Relcache / specCache / RelxillCache, src/Relcache.cpp,
 * src/Relbase.cpp:143-168, src/Relxill.cpp:296-300,405, as a device-resident re-use of the previous run):
 * after a run the scratch arena still holds the batch's intermediates.  Update the parameters (same model,
 * same n_vec) and/or the energy grid in place and run again: a vector whose whole parameter set is unchanged
 * (z aside) only repeats the final rebin, one whose relativistic parameters are unchanged keeps its line
 * profiles and emission-angle distribution and repeats only the xillver half.  The results are bit-identical
 * to a fresh evaluation.  State survives while no other batch has used the arena in between and the batch fits the
 * arena (n_vec <= the chunk capacity, 4096 by default; RELXILL_B200_CHUNK) — also when a host-buffer call cuts it into
 * pipelined pieces, each piece keeps its own slice of the arena.  relxill_batch_eval and the lmod* symbols keep their
 * last batch alive for the same purpose (an XSPEC fit varies one parameter at a time; an MCMC step leaves the rejected
 * walkers where they were). */
int relxill_b200_update_params(relxill_b200_batch *b, const double *params /* [n_vec][npar] */);
int relxill_b200_update_energy(relxill_b200_batch *b, const double *energy, int n_flux);
/* vectors of the last run that were {recomputed, re-used the relativistic half, re-used everything} */
int relxill_b200_reuse_counts(relxill_b200_batch *b, long *out3);
/* the same counts for the batch(es) the last relxill_batch_eval / lmod* call retained (summed over the devices) */
int relxill_b200_last_eval_reuse(long *out3);
/* switch the re-use off (0) / on (1, default); measurements of the full path switch it off */
void relxill_b200_set_cache(int on);
/* per-vector status after prepare/run (host copy), length n_vec */
int relxill_b200_batch_status(relxill_b200_batch *b, int *status);
void relxill_b200_free_batch(relxill_b200_batch *b);

/* ---------------------------------------------------------------- measurement support */
/* Exact algorithmic byte count of the last run of `b`: distinct xillver corner
 * rows per vector (U, host-counted from the zone indices the kernels produced) and the other
 * table/IO terms.  out[0]=bytes total, out[1]=sum of U over vectors, out[2]=xillver bytes,
 * out[3]=upper bound without cross-zone sharing, out[4]=bytes of the per-zone line profiles actually
 * produced (first to last non-zero bin of every zone), out[5]=values per zone spectrum row as filed by
 * k_xill and read by k_conv, out[6]=bytes of the xillver rows that ANY vector of the batch reads, each counted once
 * (what has to leave DRAM per launch at least; vectors share rows through L2), out[7] reserved (0). */
int relxill_b200_algorithmic_bytes(relxill_b200_batch *b, double *out8);
/* Number of kernel launches issued by the last relxill_b200_run of `b`. */
long relxill_b200_last_launches(relxill_b200_batch *b);
/* Time (ms, CUDA events on the run's stream) spent in each kernel family during the last run when
 * profiling is enabled with relxill_b200_set_profiling(1); names[] receives static strings.
 * Returns the number of entries written (<= max). */
void relxill_b200_set_profiling(int on);
int relxill_b200_kernel_times(relxill_b200_batch *b, const char **names, double *ms, long *launches, int max);

/* FP64 roofline denominator of the current device, measured on the spot: TFLOP/s of a DFMA microkernel (independent
 * chains, no memory traffic, CUDA events, best of 5 after a warm-up launch).  < 0 on error. */
double relxill_b200_measure_fp64_peak(void);

/* Grid on which the per-zone xillver spectra are blended: 1 (default) = the convolution grid — every table
 * row is rebinned once at load (the reference rebins every zone spectrum of every evaluation,
 * _rebin_spectrum, src/Relxill.cpp:461-463; the map is linear, so it commutes with the interpolation) and
 * kept in fp64 next to the fp32 table; 0 = the table grid, each zone spectrum rebinned inside k_conv.
 * Same results to rounding.  The environment variable RELXILL_B200_XILL_GRID=table, read at
 * initialisation, selects 0 and skips building the copy (1.7x the table's bytes). */
void relxill_b200_set_xill_grid(int conv_grid);
int relxill_b200_get_xill_grid(void);
/* Test hook: evaluate with the instantiation of the xillver blend kernel that takes the table's row length and
 * inclination count at run time (what a table with another energy grid gets) even on the standard 2999-bin
 * tables, whose strides are otherwise compile-time constants. */
void relxill_b200_set_xill_generic(int on);

/* Keep the intermediates that only the probes read (the fine emission-angle tables are otherwise not stored
 * unless a limb law needs them).  Off by default. */
void relxill_b200_keep_intermediates(int on);

/* Stage probes for the parity tests (device -> host copies of intermediates of vector `iv` of
 * the last run; sizes as in oracle/relxill_oracle.h).  Return 0 on success. */
int relxill_b200_probe(relxill_b200_batch *b, long iv, const char *what, double *out, long max_len);

#ifdef __cplusplus
}
#endif
#endif
