/* cfitsio entry points used by the reference, on top of minifits.h.
 * I/O only: big-endian cells -> host types, same (row, element, count)
 * addressing as cfitsio's fits_read_col.  Test infrastructure. */
#include "fitsio.h"
#include "../../relxill_b200/csrc/minifits.h"

struct oracle_fitsfile {
  mf_file *f;
  int cur; /* 1-based HDU number */
};

#define ERR_OPEN 104
#define ERR_HDU 301
#define ERR_COL 219
#define ERR_READ 108

int fits_open_table(fitsfile **fptr, const char *filename, int iomode, int *status) {
  (void) iomode;
  if (*status) return *status;
  mf_file *f = mf_open(filename);
  if (!f) { *fptr = NULL; return (*status = ERR_OPEN); }
  fitsfile *p = (fitsfile *) calloc(1, sizeof(*p));
  p->f = f;
  p->cur = 0;
  for (int i = 0; i < f->nhdu; i++)
    if (f->hdus[i].is_table) { p->cur = i + 1; break; }
  if (!p->cur) { mf_close(f); free(p); *fptr = NULL; return (*status = ERR_HDU); }
  *fptr = p;
  return 0;
}

int fits_close_file(fitsfile *fptr, int *status) {
  if (fptr) { mf_close(fptr->f); free(fptr); }
  return *status;
}

int fits_movnam_hdu(fitsfile *fptr, int hdutype, const char *extname, int extver, int *status) {
  (void) hdutype; (void) extver;
  if (*status) return *status;
  int h = mf_find_hdu(fptr->f, extname);
  if (!h) return (*status = ERR_HDU);
  fptr->cur = h;
  return 0;
}

int fits_movabs_hdu(fitsfile *fptr, int hdunum, int *exttype, int *status) {
  if (*status) return *status;
  if (hdunum < 1 || hdunum > fptr->f->nhdu) return (*status = ERR_HDU);
  fptr->cur = hdunum;
  if (exttype) *exttype = fptr->f->hdus[hdunum - 1].is_table ? BINARY_TBL : 0;
  return 0;
}

int fits_get_num_rows(fitsfile *fptr, long *nrows, int *status) {
  if (*status) return *status;
  *nrows = fptr->f->hdus[fptr->cur - 1].nrows;
  return 0;
}

int fits_get_colnum(fitsfile *fptr, int casesen, const char *templt, int *colnum, int *status) {
  (void) casesen;
  if (*status) return *status;
  int c = mf_find_col(&fptr->f->hdus[fptr->cur - 1], templt);
  if (!c) return (*status = ERR_COL);
  *colnum = c;
  return 0;
}

int fits_read_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                  LONGLONG nelem, void *nulval, void *array, int *anynul, int *status) {
  (void) nulval;
  if (*status) return *status;
  if (anynul) *anynul = 0;
  const mf_hdu *h = &fptr->f->hdus[fptr->cur - 1];
  int rc;
  if (datatype == TSTRING) {
    char **dst = (char **) array;
    rc = 0;
    for (LONGLONG k = 0; k < nelem && !rc; k++) rc = mf_read_str(h, colnum, (long) (firstrow + k), dst[k], 8);
  } else {
    char kind = datatype == TFLOAT ? 'f' : datatype == TDOUBLE ? 'd' : 'i';
    rc = mf_read(h, colnum, (long) firstrow, (long) firstelem, (long) nelem, kind, array);
  }
  if (rc) return (*status = ERR_READ);
  return 0;
}

void fits_get_errstatus(int status, char *errtext) { snprintf(errtext, 30, "minifits shim error %d", status); }

int fits_create_file(fitsfile **fptr, const char *filename, int *status) {
  (void) filename; *fptr = NULL; return (*status = ERR_OPEN);
}
int fits_create_tbl(fitsfile *fptr, int tbltype, LONGLONG naxis2, int tfields, char **ttype, char **tform,
                    char **tunit, const char *extname, int *status) {
  (void) fptr; (void) tbltype; (void) naxis2; (void) tfields; (void) ttype; (void) tform; (void) tunit; (void) extname;
  return (*status = ERR_OPEN);
}
int fits_write_col(fitsfile *fptr, int datatype, int colnum, LONGLONG firstrow, LONGLONG firstelem,
                   LONGLONG nelem, void *array, int *status) {
  (void) fptr; (void) datatype; (void) colnum; (void) firstrow; (void) firstelem; (void) nelem; (void) array;
  return (*status = ERR_OPEN);
}
int fits_write_key(fitsfile *fptr, int datatype, const char *keyname, void *value, const char *comm,
                   int *status) {
  (void) fptr; (void) datatype; (void) keyname; (void) value; (void) comm;
  return (*status = ERR_OPEN);
}
