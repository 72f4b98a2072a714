// stub of png.h (declarations only): lets RELION's jaz/gravis headers parse without libpng installed
#pragma once
#include <stdio.h>
#include <stddef.h>
#include <setjmp.h>
extern "C" {
typedef unsigned char png_byte; typedef png_byte *png_bytep; typedef png_byte **png_bytepp; typedef unsigned int png_uint_32;
typedef struct png_struct_def png_struct; typedef png_struct *png_structp; typedef png_struct **png_structpp;
typedef struct png_info_def png_info; typedef png_info *png_infop; typedef png_info **png_infopp;
typedef void *png_voidp; typedef const char *png_const_charp; typedef void (*png_error_ptr)(png_structp, png_const_charp);
typedef size_t png_size_t; typedef size_t png_alloc_size_t;
#define PNG_LIBPNG_VER_STRING "stub"
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_PALETTE 3
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_RGB_ALPHA 6
#define PNG_COLOR_TYPE_GRAY_ALPHA 4
#define PNG_COMPRESSION_TYPE_BASE 0
#define PNG_FILTER_TYPE_BASE 0
#define PNG_INTERLACE_NONE 0
#define PNG_FILLER_AFTER 1
png_structp png_create_read_struct(png_const_charp, png_voidp, png_error_ptr, png_error_ptr);
png_structp png_create_write_struct(png_const_charp, png_voidp, png_error_ptr, png_error_ptr);
png_infop png_create_info_struct(png_structp);
void png_destroy_read_struct(png_structpp, png_infopp, png_infopp);
void png_destroy_write_struct(png_structpp, png_infopp);
void png_init_io(png_structp, FILE *);
void png_read_info(png_structp, png_infop); void png_read_update_info(png_structp, png_infop);
void png_read_image(png_structp, png_bytepp); void png_read_end(png_structp, png_infop);
void png_write_info(png_structp, png_infop); void png_write_image(png_structp, png_bytepp); void png_write_end(png_structp, png_infop);
png_uint_32 png_get_IHDR(png_structp, png_infop, png_uint_32 *, png_uint_32 *, int *, int *, int *, int *, int *);
void png_set_IHDR(png_structp, png_infop, png_uint_32, png_uint_32, int, int, int, int, int);
png_size_t png_get_rowbytes(png_structp, png_infop);
void png_set_expand_gray_1_2_4_to_8(png_structp); void png_set_filler(png_structp, png_uint_32, int); void png_set_gray_to_rgb(png_structp);
int png_set_interlace_handling(png_structp); void png_set_palette_to_rgb(png_structp); void png_set_strip_16(png_structp);
int png_sig_cmp(png_bytep, png_size_t, png_size_t);
png_voidp png_malloc(png_structp, png_alloc_size_t); void png_free(png_structp, png_voidp);
jmp_buf *png_set_longjmp_fn(png_structp, void (*)(jmp_buf, int), size_t);
#define png_jmpbuf(png_ptr) (*png_set_longjmp_fn((png_ptr), longjmp, sizeof(jmp_buf)))
}
