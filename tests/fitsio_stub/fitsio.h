/* Declarations-only stand-in for cfitsio's fitsio.h (absent from this image), used by the tests for a
 * -fsyntax-only check that the reference's FITS-writing applications compile unchanged against the kerrtrace
 * shim headers.  Nothing here is ever linked or executed. */
#ifndef KT_TEST_FITSIO_STUB_H
#define KT_TEST_FITSIO_STUB_H
#include <stdio.h>
typedef struct fitsfile_stub fitsfile;
enum { BYTE_IMG = 8, SHORT_IMG = 16, LONG_IMG = 32, FLOAT_IMG = -32, DOUBLE_IMG = -64 };
enum { TLOGICAL = 14, TSTRING = 16, TINT = 31, TLONG = 41, TFLOAT = 42, TDOUBLE = 82 };
enum { BINARY_TBL = 2 };
#ifdef __cplusplus
extern "C" {
#endif
void fits_report_error(FILE*, int);
int fits_create_file(fitsfile**, const char*, int*);
int fits_close_file(fitsfile*, int*);
int fits_create_img(fitsfile*, int, int, long*, int*);
int fits_write_pix(fitsfile*, int, long*, long long, void*, int*);
int fits_create_tbl(fitsfile*, int, long long, int, char**, char**, char**, const char*, int*);
int fits_write_col(fitsfile*, int, int, long long, long long, long long, void*, int*);
int fits_write_key(fitsfile*, int, const char*, void*, const char*, int*);
int fits_write_comment(fitsfile*, const char*, int*);
#ifdef __cplusplus
}
#endif
#endif
