/* minifits.c -- see fitsio.h.  Each HDU is buffered in memory (header cards + data) and written when the next HDU
 * is created or the file is closed, because the applications add keywords after the pixel data. */
#include "fitsio.h"

#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct minifits_file {
    FILE* fp;
    int nhdu;          /* HDUs started so far */
    char* cards;       /* current HDU: 80-byte cards */
    size_t ncards, cap_cards;
    unsigned char* data;
    size_t ndata;      /* bytes of data in the current HDU */
    int bitpix;
    long long npix;
    int open_hdu;
};

static void add_card(fitsfile* f, const char* text)
{
    if (f->ncards == f->cap_cards) {
        f->cap_cards = f->cap_cards ? 2 * f->cap_cards : 64;
        f->cards = (char*)realloc(f->cards, f->cap_cards * 80);
    }
    char* c = f->cards + 80 * f->ncards++;
    memset(c, ' ', 80);
    size_t n = strlen(text);
    memcpy(c, text, n > 80 ? 80 : n);
}

static void add_value_card(fitsfile* f, const char* key, const char* value, const char* comment)
{
    /* fixed format: KEYWORD = <value right-justified to column 30> / comment ; HIERARCH for non-standard names */
    char card[256];
    int standard = strlen(key) <= 8;
    for (const char* p = key; *p && standard; p++)
        if (!(isupper((unsigned char)*p) || isdigit((unsigned char)*p) || *p == '-' || *p == '_')) standard = 0;
    if (standard) {
        if (value[0] == '\'') snprintf(card, sizeof card, "%-8s= %-20s", key, value);
        else snprintf(card, sizeof card, "%-8s= %20s", key, value);
    } else {
        snprintf(card, sizeof card, "HIERARCH %s = %s", key, value);
    }
    if (comment && *comment) {
        size_t n = strlen(card);
        snprintf(card + n, sizeof card - n, " / %s", comment);
    }
    add_card(f, card);
}

static int flush_hdu(fitsfile* f)
{
    if (!f->open_hdu) return 0;
    static const char blank[2880] = {0};
    add_card(f, "END");
    size_t hbytes = f->ncards * 80;
    if (fwrite(f->cards, 1, hbytes, f->fp) != hbytes) return MINIFITS_WRITE_ERROR;
    size_t pad = (2880 - hbytes % 2880) % 2880;
    for (size_t i = 0; i < pad; i++) fputc(' ', f->fp);
    if (f->ndata) {
        if (fwrite(f->data, 1, f->ndata, f->fp) != f->ndata) return MINIFITS_WRITE_ERROR;
        pad = (2880 - f->ndata % 2880) % 2880;
        if (pad && fwrite(blank, 1, pad, f->fp) != pad) return MINIFITS_WRITE_ERROR;
    }
    f->ncards = 0;
    free(f->data);
    f->data = NULL;
    f->ndata = 0;
    f->open_hdu = 0;
    return 0;
}

void fits_report_error(FILE* stream, int status)
{
    if (status) fprintf(stream, "minifits error status %d\n", status);
}

int fits_create_file(fitsfile** fptr, const char* filename, int* status)
{
    if (*status > 0) return *status;
    const char* name = filename;
    if (name[0] == '!') name++;       /* cfitsio: '!' = clobber; we always overwrite */
    fitsfile* f = (fitsfile*)calloc(1, sizeof(fitsfile));
    f->fp = fopen(name, "wb");
    if (!f->fp) {
        free(f);
        return *status = MINIFITS_FILE_NOT_CREATED;
    }
    *fptr = f;
    return 0;
}

int fits_close_file(fitsfile* f, int* status)
{
    int rc = flush_hdu(f);
    if (fclose(f->fp) != 0 && !rc) rc = MINIFITS_WRITE_ERROR;
    free(f->cards);
    free(f);
    if (rc && *status <= 0) *status = rc;
    return *status;
}

int fits_create_img(fitsfile* f, int bitpix, int naxis, long* naxes, int* status)
{
    if (*status > 0) return *status;
    if ((*status = flush_hdu(f)) != 0) return *status;
    char v[64];
    if (f->nhdu == 0) add_value_card(f, "SIMPLE", "T", "file does conform to FITS standard");
    else add_value_card(f, "XTENSION", "'IMAGE   '", "IMAGE extension");
    snprintf(v, sizeof v, "%d", bitpix);
    add_value_card(f, "BITPIX", v, "number of bits per data pixel");
    snprintf(v, sizeof v, "%d", naxis);
    add_value_card(f, "NAXIS", v, "number of data axes");
    f->npix = naxis > 0 ? 1 : 0;
    for (int i = 0; i < naxis; i++) {
        char k[16];
        snprintf(k, sizeof k, "NAXIS%d", i + 1);
        snprintf(v, sizeof v, "%ld", naxes[i]);
        add_value_card(f, k, v, "length of data axis");
        f->npix *= naxes[i];
    }
    if (f->nhdu == 0) add_value_card(f, "EXTEND", "T", "FITS dataset may contain extensions");
    else {
        add_value_card(f, "PCOUNT", "0", "required keyword; must = 0");
        add_value_card(f, "GCOUNT", "1", "required keyword; must = 1");
    }
    f->bitpix = bitpix;
    f->ndata = (size_t)f->npix * (size_t)(abs(bitpix) / 8);
    f->data = f->ndata ? (unsigned char*)calloc(f->ndata, 1) : NULL;
    f->nhdu++;
    f->open_hdu = 1;
    return 0;
}

static void put_be(unsigned char* dst, const void* src, int nbytes)
{
    const unsigned char* s = (const unsigned char*)src;
    for (int i = 0; i < nbytes; i++) dst[i] = s[nbytes - 1 - i];   /* host is little-endian x86 / arm64 */
}

int fits_write_pix(fitsfile* f, int datatype, long* firstpix, long long nelem, void* array, int* status)
{
    if (*status > 0) return *status;
    if (!f->open_hdu || !f->data) return *status = MINIFITS_NOT_IMAGE;
    (void)firstpix;   /* the reference always starts at pixel (1,1) */
    if (nelem > f->npix) nelem = f->npix;
    const int width = abs(f->bitpix) / 8;
    for (long long i = 0; i < nelem; i++) {
        double v;
        if (datatype == TDOUBLE) v = ((const double*)array)[i];
        else if (datatype == TFLOAT) v = ((const float*)array)[i];
        else if (datatype == TINT) v = ((const int*)array)[i];
        else if (datatype == TLONG) v = (double)((const long*)array)[i];
        else return *status = MINIFITS_BAD_DATATYPE;
        unsigned char* dst = f->data + (size_t)i * width;
        if (f->bitpix == DOUBLE_IMG) { put_be(dst, &v, 8); }
        else if (f->bitpix == FLOAT_IMG) { float x = (float)v; put_be(dst, &x, 4); }
        else if (f->bitpix == LONG_IMG) { int32_t x = (int32_t)v; put_be(dst, &x, 4); }
        else if (f->bitpix == SHORT_IMG) { int16_t x = (int16_t)v; put_be(dst, &x, 2); }
        else { dst[0] = (unsigned char)v; }
    }
    return 0;
}

int fits_create_tbl(fitsfile* f, int tbltype, long long naxis2, int tfields, char** ttype, char** tform, char** tunit,
                    const char* extname, int* status)
{
    (void)f; (void)tbltype; (void)naxis2; (void)tfields; (void)ttype; (void)tform; (void)tunit; (void)extname;
    if (*status > 0) return *status;
    return *status = MINIFITS_NOT_TABLE;   /* binary tables are outside the scope of this writer */
}

int fits_write_col(fitsfile* f, int datatype, int colnum, long long firstrow, long long firstelem, long long nelem,
                   void* array, int* status)
{
    (void)f; (void)datatype; (void)colnum; (void)firstrow; (void)firstelem; (void)nelem; (void)array;
    if (*status > 0) return *status;
    return *status = MINIFITS_NOT_TABLE;
}

int fits_write_key(fitsfile* f, int datatype, const char* keyname, void* value, const char* comment, int* status)
{
    if (*status > 0) return *status;
    if (!f->open_hdu) return *status = MINIFITS_NOT_IMAGE;
    char v[128];
    if (datatype == TINT) snprintf(v, sizeof v, "%d", *(const int*)value);
    else if (datatype == TLONG) snprintf(v, sizeof v, "%ld", *(const long*)value);
    else if (datatype == TDOUBLE) {
        snprintf(v, sizeof v, "%.15G", *(const double*)value);
        if (!strpbrk(v, ".EN")) strncat(v, ".", sizeof v - strlen(v) - 1);      /* keep it a floating-point literal */
    }
    else if (datatype == TFLOAT) snprintf(v, sizeof v, "%.7G", (double)*(const float*)value);
    else if (datatype == TLOGICAL) snprintf(v, sizeof v, "%s", *(const int*)value ? "T" : "F");
    else if (datatype == TSTRING) {
        /* quoted, single quotes doubled, at least 8 characters between the quotes */
        const char* s = (const char*)value;
        size_t n = 0;
        v[n++] = '\'';
        for (; *s && n < sizeof v - 12; s++) {
            if (*s == '\'') v[n++] = '\'';
            v[n++] = *s;
        }
        while (n < 9) v[n++] = ' ';
        v[n++] = '\'';
        v[n] = 0;
    } else return *status = MINIFITS_BAD_DATATYPE;
    add_value_card(f, keyname, v, comment);
    return 0;
}

int fits_write_comment(fitsfile* f, const char* comment, int* status)
{
    if (*status > 0) return *status;
    if (!f->open_hdu) return *status = MINIFITS_NOT_IMAGE;
    char card[96];
    snprintf(card, sizeof card, "COMMENT %.72s", comment);
    add_card(f, card);
    return 0;
}
