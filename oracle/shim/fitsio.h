/* fitsio.h shim — declarations only, for compiling the unmodified reference
 * sources in this image (no cfitsio installed).  The twelve entry points the
 * reference links against are implemented in oracle/shim/fits_shim.c on top
 * of relxill_b200/csrc/minifits.h.  Test infrastructure, not product code. */
#ifndef ORACLE_SHIM_FITSIO_H_
#define ORACLE_SHIM_FITSIO_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_fitsfile fitsfile;
typedef long long LONGLONG;

#define READONLY 0
#define BINARY_TBL 2
#define CASEINSEN 0
#define TSTRING 16
#define TINT 31
#define TFLOAT 42
#define TDOUBLE 82

int fits_open_table(fitsfile **fptr, const char *filename, int iomode, int *status);
int fits_close_file(fitsfile *fptr, int *status);
int fits_movnam_hdu(fitsfile *fptr, int hdutype, const char *extname, int extver, int *status);
int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *exttype, int *status);
int fits_get_num_rows(fitsfile *fptr, long *nrows, int *status);
int fits_get_colnum(fitsfile *fptr, int casesen, const char *templt, int *colnum, int *status);
int fits_read_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                  LONGLONG nelem, void *nulval, void *array, int *anynul, int *status);
void fits_get_errstatus(int status, char *errtext);

/* only referenced by the out-of-scope relxillBB debug writer: stubs */
int fits_create_file(fitsfile **fptr, const char *filename, int *status);
int fits_create_tbl(fitsfile *fptr, int tbltype, LONGLONG naxis2, int tfields, char **ttype, char **tform,
                    char **tunit, const char *extname, int *status);
int fits_write_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                   LONGLONG nelem, void *array, int *status);
int fits_write_key(fitsfile *fptr, int datatype, const char *keyname, void *value, const char *comm,
                   int *status);

#ifdef __cplusplus
}
#endif
#endif
