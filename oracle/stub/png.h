/* TEST INFRASTRUCTURE - not part of the product.
 *
 * Minimal stand-in for libpng's <png.h>, which is not installed in this image.
 * The reference's pngloss_image.c includes <png.h> only for the five per-row
 * filter masks it writes into row_filters[] (reference src/pngloss_image.c:18,
 * :290-306).  These values are the public libpng 1.x ABI constants.
 */
#ifndef PNGLOSS_ORACLE_STUB_PNG_H
#define PNGLOSS_ORACLE_STUB_PNG_H
#define PNG_FILTER_NONE  0x08
#define PNG_FILTER_SUB   0x10
#define PNG_FILTER_UP    0x20
#define PNG_FILTER_AVG   0x40
#define PNG_FILTER_PAETH 0x80
#endif
