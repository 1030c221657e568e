/*
 * minifits -- a minimal, dependency-free writer for the nine cfitsio calls the reference's include/fits_output.h makes
 * (fits_create_file, fits_close_file, fits_create_img, fits_write_pix, fits_create_tbl, fits_write_col, fits_write_key,
 * fits_write_comment, fits_report_error), so that the FITS-writing applications (imageplane_disc_image*, caustic_*) link
 * and run where cfitsio is not installed.  It writes standard-conforming FITS: a primary
 * HDU plus IMAGE extensions (BITPIX 8 / 16 / 32 / -32 / -64), 2880-byte blocks, big-endian data, fixed-format keyword cards
 * (HIERARCH convention for keyword names that are not 8-character upper-case FITS names, as cfitsio does).
 * Binary tables are not implemented (no application in scope writes one): fits_create_tbl reports an error status.
 * Host-side output only; nothing here is on the hot path.
 */
#ifndef KT_MINIFITS_FITSIO_H
#define KT_MINIFITS_FITSIO_H
#include <stdio.h>

typedef struct minifits_file fitsfile;

#define BYTE_IMG 8
#define SHORT_IMG 16
#define LONG_IMG 32
#define FLOAT_IMG (-32)
#define DOUBLE_IMG (-64)

#define TLOGICAL 14
#define TSTRING 16
#define TINT 31
#define TLONG 41
#define TFLOAT 42
#define TDOUBLE 82

#define BINARY_TBL 2

#define MINIFITS_FILE_NOT_CREATED 105
#define MINIFITS_WRITE_ERROR 106
#define MINIFITS_NOT_IMAGE 233
#define MINIFITS_BAD_DATATYPE 410
#define MINIFITS_NOT_TABLE 235

#ifdef __cplusplus
extern "C" {
#endif
void fits_report_error(FILE* stream, int status);
int fits_create_file(fitsfile** fptr, const char* filename, int* status);
int fits_close_file(fitsfile* fptr, int* status);
int fits_create_img(fitsfile* fptr, int bitpix, int naxis, long* naxes, int* status);
int fits_write_pix(fitsfile* fptr, int datatype, long* firstpix, long long nelem, void* array, int* status);
int fits_create_tbl(fitsfile* fptr, int tbltype, long long naxis2, int tfields, char** ttype, char** tform, char** tunit,
                    const char* extname, int* status);
int fits_write_col(fitsfile* fptr, int datatype, int colnum, long long firstrow, long long firstelem, long long nelem,
                   void* array, int* status);
int fits_write_key(fitsfile* fptr, int datatype, const char* keyname, void* value, const char* comment, int* status);
int fits_write_comment(fitsfile* fptr, const char* comment, int* status);
#ifdef __cplusplus
}
#endif
#endif
