/* minifits.h — read-only FITS binary-table access (header-only, C99 / C++).
 *
 * Just enough of FITS to read the relxill tables: mmap the file, index the
 * HDUs, locate BINTABLE columns and copy big-endian cells out as float / double /
 * int / string.  Every TFORM code of the FITS standard gets its true byte width, so
 * that a column this reader cannot decode (bit, complex, variable-length P/Q) never
 * shifts the columns behind it: reading such a column fails, reading its neighbours
 * works.  A table whose column widths do not add up to NAXIS1 is rejected as a whole
 * (every read fails).  TSCALn / TZEROn are applied like cfitsio does.  No cfitsio in this
 * image, so both the CUDA library's table loader and the oracle's cfitsio shim
 * sit on this reader (I/O only, no arithmetic).
 */
#ifndef RELXILL_B200_MINIFITS_H_
#define RELXILL_B200_MINIFITS_H_

#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define MF_BLOCK 2880
#define MF_MAXCOLS 1024

typedef struct {
  char name[72];
  char code;       /* TFORM data type: 'E','D','J','I','K','B','L' numeric; 'A' string; others: not decodable */
  long width;      /* repeat count */
  long offset;     /* byte offset inside a row */
  int elsize;      /* bytes per element (0: unknown code, the whole table is rejected) */
  double tscal, tzero;
} mf_col;

typedef struct {
  char extname[72];
  int is_table;
  long nrows, rowbytes;
  int ncols;
  int bad_layout;  /* column widths do not add up to NAXIS1 / unknown TFORM code: reads fail */
  mf_col *cols;
  const unsigned char *data;
} mf_hdu;

/* bytes per element of a TFORM data-type code (FITS standard 4.0, table 18); 0 = not a valid code */
static inline int mf__elsize(char code) {
  switch (code) {
    case 'L': case 'B': case 'A': return 1;
    case 'I': return 2;
    case 'J': case 'E': return 4;
    case 'K': case 'D': case 'C': case 'P': return 8;
    case 'M': case 'Q': return 16;
    case 'X': return -1;   /* bits: ceil(repeat / 8) bytes for the whole cell */
    default: return 0;
  }
}

typedef struct {
  const unsigned char *base;
  size_t size;
  int nhdu;
  mf_hdu *hdus;
} mf_file;

static inline int mf__card_value(const char *card, char *out, size_t n) {
  /* value part of "KEY     = value / comment" with quotes stripped */
  if (card[8] != '=') return 0;
  const char *p = card + 10;
  const char *end = card + 80;
  while (p < end && *p == ' ') p++;
  size_t k = 0;
  if (p < end && *p == '\'') {
    p++;
    while (p < end && *p != '\'' && k + 1 < n) out[k++] = *p++;
    while (k > 0 && out[k - 1] == ' ') k--;
  } else {
    while (p < end && *p != ' ' && *p != '/' && k + 1 < n) out[k++] = *p++;
  }
  out[k] = 0;
  return 1;
}

static inline void mf_close(mf_file *f) {
  if (!f) return;
  for (int i = 0; i < f->nhdu; i++) free(f->hdus[i].cols);
  free(f->hdus);
  if (f->base) munmap((void *) f->base, f->size);
  free(f);
}

static inline mf_file *mf_open(const char *path) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return NULL;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < MF_BLOCK) { close(fd); return NULL; }
  void *m = mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return NULL;
  mf_file *f = (mf_file *) calloc(1, sizeof(mf_file));
  f->base = (const unsigned char *) m;
  f->size = (size_t) st.st_size;
  if (memcmp(f->base, "SIMPLE", 6) != 0) { mf_close(f); return NULL; }
  int cap = 16;
  f->hdus = (mf_hdu *) calloc((size_t) cap, sizeof(mf_hdu));
  size_t pos = 0;
  while (pos + MF_BLOCK <= f->size) {
    if (f->nhdu == cap) {
      cap *= 2;
      f->hdus = (mf_hdu *) realloc(f->hdus, (size_t) cap * sizeof(mf_hdu));
      memset(f->hdus + f->nhdu, 0, (size_t) (cap - f->nhdu) * sizeof(mf_hdu));
    }
    mf_hdu *h = &f->hdus[f->nhdu];
    memset(h, 0, sizeof(*h));
    long naxis = 0, naxes[8] = {0}, tfields = 0, pcount = 0;
    mf_col *cols = NULL;
    int done = 0;
    char val[80];
    while (!done && pos + MF_BLOCK <= f->size) {
      for (int i = 0; i < MF_BLOCK && !done; i += 80) {
        const char *card = (const char *) f->base + pos + i;
        if (memcmp(card, "END     ", 8) == 0) { done = 1; break; }
        if (!mf__card_value(card, val, sizeof(val))) continue;
        if (memcmp(card, "XTENSION", 8) == 0) h->is_table = (strcmp(val, "BINTABLE") == 0);
        else if (memcmp(card, "NAXIS   ", 8) == 0) naxis = atol(val);
        else if (memcmp(card, "NAXIS", 5) == 0 && card[5] >= '1' && card[5] <= '8' && card[6] == ' ')
          naxes[card[5] - '1'] = atol(val);
        else if (memcmp(card, "PCOUNT  ", 8) == 0) pcount = atol(val);
        else if (memcmp(card, "EXTNAME ", 8) == 0) { strncpy(h->extname, val, sizeof(h->extname) - 1); }
        else if (memcmp(card, "TFIELDS ", 8) == 0) {
          tfields = atol(val);
          if (tfields > MF_MAXCOLS) tfields = MF_MAXCOLS;
          cols = (mf_col *) calloc((size_t) (tfields > 0 ? tfields : 1), sizeof(mf_col));
          for (long c = 0; c < tfields; c++) cols[c].tscal = 1.0;
        } else if (cols && (memcmp(card, "TTYPE", 5) == 0 || memcmp(card, "TFORM", 5) == 0)) {
          long idx = atol(card + 5);
          if (idx >= 1 && idx <= tfields) {
            if (card[1] == 'T') {
              strncpy(cols[idx - 1].name, val, sizeof(cols[idx - 1].name) - 1);
            } else {
              char *e = NULL;
              long rep = strtol(val, &e, 10);
              if (e == val) rep = 1;
              cols[idx - 1].width = rep;
              cols[idx - 1].code = *e;
              cols[idx - 1].elsize = mf__elsize(*e);
            }
          }
        } else if (cols && (memcmp(card, "TSCAL", 5) == 0 || memcmp(card, "TZERO", 5) == 0)) {
          long idx = atol(card + 5);
          if (idx >= 1 && idx <= tfields) {
            if (card[1] == 'S') cols[idx - 1].tscal = atof(val);
            else cols[idx - 1].tzero = atof(val);
          }
        }
      }
      pos += MF_BLOCK;
    }
    size_t dsize = 0;
    if (naxis > 0) {
      dsize = 1;
      for (long a = 0; a < naxis; a++) dsize *= (size_t) naxes[a];
      dsize += (size_t) pcount;
    }
    if (h->is_table) {
      h->rowbytes = naxes[0];
      h->nrows = naxes[1];
      h->ncols = (int) tfields;
      h->cols = cols;
      long off = 0;
      for (int c = 0; c < h->ncols; c++) {
        cols[c].offset = off;
        if (cols[c].elsize == 0) h->bad_layout = 1;
        off += (cols[c].elsize < 0) ? (cols[c].width + 7) / 8 : cols[c].width * cols[c].elsize;
      }
      if (off != h->rowbytes) h->bad_layout = 1;
      h->data = f->base + pos;
    } else {
      free(cols);
    }
    f->nhdu++;
    pos += dsize + ((MF_BLOCK - dsize % MF_BLOCK) % MF_BLOCK);
  }
  return f;
}

/* 1-based HDU number of the first table named `extname` (case-insensitive), 0 if absent */
static inline int mf_find_hdu(const mf_file *f, const char *extname) {
  for (int i = 0; i < f->nhdu; i++)
    if (f->hdus[i].is_table && strcasecmp(f->hdus[i].extname, extname) == 0) return i + 1;
  return 0;
}

/* 1-based column number, 0 if absent */
static inline int mf_find_col(const mf_hdu *h, const char *name) {
  for (int c = 0; c < h->ncols; c++)
    if (strcasecmp(h->cols[c].name, name) == 0) return c + 1;
  return 0;
}

static inline int mf__numeric(char code) {
  return code == 'E' || code == 'D' || code == 'J' || code == 'I' || code == 'K' || code == 'B' || code == 'L';
}

static inline double mf__cell(const mf_col *c, const unsigned char *p) {
  if (c->code == 'I') return (double) (int16_t) (((uint16_t) p[0] << 8) | p[1]);
  if (c->code == 'B') return (double) p[0];
  if (c->code == 'L') return (p[0] == 'T') ? 1.0 : 0.0;
  if (c->code == 'K') {
    uint64_t u = 0;
    for (int i = 0; i < 8; i++) u = (u << 8) | p[i];
    return (double) (int64_t) u;
  }
  if (c->code == 'E') {
    uint32_t u = ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3];
    float v;
    memcpy(&v, &u, 4);
    return (double) v;
  } else if (c->code == 'D') {
    uint64_t u = 0;
    for (int i = 0; i < 8; i++) u = (u << 8) | p[i];
    double v;
    memcpy(&v, &u, 8);
    return v;
  } else { /* 'J' */
    uint32_t u = ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3];
    return (double) (int32_t) u;
  }
}

/* cfitsio-style read: `nelem` cells starting at (firstrow, firstelem), both
 * 1-based, running on into the following rows.  out_kind: 'f','d','i'. */
static inline int mf_read(const mf_hdu *h, int colnum, long firstrow, long firstelem, long nelem,
                          char out_kind, void *out) {
  if (colnum < 1 || colnum > h->ncols || h->bad_layout) return 1;
  const mf_col *c = &h->cols[colnum - 1];
  if (!mf__numeric(c->code) || c->width < 1 || firstrow < 1 || firstelem < 1) return 1;
  long idx = (firstrow - 1) * c->width + (firstelem - 1);
  for (long k = 0; k < nelem; k++, idx++) {
    long row = idx / c->width, el = idx % c->width;
    if (row >= h->nrows) return 2;
    double v = mf__cell(c, h->data + row * h->rowbytes + c->offset + el * c->elsize);
    if (c->tscal != 1.0 || c->tzero != 0.0) v = v * c->tscal + c->tzero;
    if (out_kind == 'f') ((float *) out)[k] = (float) v;
    else if (out_kind == 'd') ((double *) out)[k] = v;
    else ((int *) out)[k] = (int) v;
  }
  return 0;
}

/* string cell of row `row` (1-based), trailing blanks stripped, at most n-1 chars */
static inline int mf_read_str(const mf_hdu *h, int colnum, long row, char *out, size_t n) {
  if (colnum < 1 || colnum > h->ncols || h->bad_layout) return 1;
  const mf_col *c = &h->cols[colnum - 1];
  if (c->code != 'A' || row < 1 || row > h->nrows) return 1;
  const unsigned char *p = h->data + (row - 1) * h->rowbytes + c->offset;
  size_t k = 0;
  for (; k < (size_t) c->width && k + 1 < n; k++) out[k] = (char) p[k];
  while (k > 0 && (out[k - 1] == ' ' || out[k - 1] == 0)) k--;
  out[k] = 0;
  return 0;
}

/* bulk big-endian float32 -> host float copy of a whole vector cell range (fast path for xillver rows) */
static inline void mf_copy_f32(const unsigned char *src, float *dst, long n) {
  for (long i = 0; i < n; i++) {
    uint32_t u;
    memcpy(&u, src + 4 * i, 4);
    u = __builtin_bswap32(u);
    memcpy(dst + i, &u, 4);
  }
}

#endif /* RELXILL_B200_MINIFITS_H_ */
