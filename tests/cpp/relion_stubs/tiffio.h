// stub of tiffio.h (declarations only): lets RELION's headers parse without libtiff installed
#pragma once
#include <stdint.h>
#include <stddef.h>
extern "C" {
typedef struct tiff TIFF;
typedef int64_t tmsize_t; typedef tmsize_t tsize_t; typedef uint64_t toff_t; typedef void *thandle_t; typedef void *tdata_t;
typedef uint32_t tstrip_t; typedef uint16_t uint16; typedef uint32_t uint32;
typedef tmsize_t (*TIFFReadWriteProc)(thandle_t, void *, tmsize_t);
typedef toff_t (*TIFFSeekProc)(thandle_t, toff_t, int);
typedef int (*TIFFCloseProc)(thandle_t);
typedef toff_t (*TIFFSizeProc)(thandle_t);
typedef int (*TIFFMapFileProc)(thandle_t, void **base, toff_t *size);
typedef void (*TIFFUnmapFileProc)(thandle_t, void *base, toff_t size);
#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_BITSPERSAMPLE 258
#define TIFFTAG_ORIENTATION 274
#define TIFFTAG_XRESOLUTION 282
#define TIFFTAG_RESOLUTIONUNIT 296
#define TIFFTAG_SAMPLEFORMAT 339
#define RESUNIT_NONE 1
#define RESUNIT_INCH 2
#define RESUNIT_CENTIMETER 3
#define SAMPLEFORMAT_UINT 1
#define SAMPLEFORMAT_INT 2
#define SAMPLEFORMAT_IEEEFP 3
#define ORIENTATION_TOPLEFT 1
#define ORIENTATION_BOTLEFT 4
TIFF *TIFFOpen(const char *, const char *);
TIFF *TIFFClientOpen(const char *, const char *, thandle_t, TIFFReadWriteProc, TIFFReadWriteProc, TIFFSeekProc, TIFFCloseProc, TIFFSizeProc, TIFFMapFileProc, TIFFUnmapFileProc);
void TIFFClose(TIFF *);
int TIFFGetField(TIFF *, uint32_t, ...);
int TIFFGetFieldDefaulted(TIFF *, uint32_t, ...);
int TIFFSetDirectory(TIFF *, uint16_t);
tmsize_t TIFFStripSize(TIFF *);
tstrip_t TIFFNumberOfStrips(TIFF *);
tmsize_t TIFFReadEncodedStrip(TIFF *, tstrip_t, void *, tmsize_t);
void *_TIFFmalloc(tmsize_t); void _TIFFfree(void *);
}
