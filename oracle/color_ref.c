/* oracle/color_ref.c -- TEST INFRASTRUCTURE ONLY.  A shim around the reference's OWN colour macros: this file contains no colour
 * arithmetic; it includes example/inc/color_util.h from the reference tree at build time (oracle/Makefile passes
 * -I$(REF)/example/inc) and applies the macros the way the reference's tools do (example/src/icer_util.c:69-121:
 * rgb888_packed_to_yuv / yuv_to_rgb888_packed with int32_t r, g, b / y, u, v).  Built into oracle/_ref/libcolor_ref.so, used by
 * tests/test_gpu_parity.py to check k_rgb_to_yuv / k_yuv_to_rgb against the compiled macros themselves. */
#include <stddef.h>
#include <stdint.h>
#include "color_util.h"

void colorref_rgb888_to_yuv(const uint8_t *rgb, uint16_t *yc, uint16_t *uc, uint16_t *vc, size_t n) {
    for (size_t i = 0; i < n; i++) {
        int32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        yc[i] = CRGB2Y(r, g, b);
        uc[i] = CRGB2Cb(r, g, b);
        vc[i] = CRGB2Cr(r, g, b);
    }
}

void colorref_yuv_to_rgb888(const uint16_t *yc, const uint16_t *uc, const uint16_t *vc, uint8_t *rgb, size_t n) {
    for (size_t i = 0; i < n; i++) {
        int32_t y = yc[i], u = uc[i], v = vc[i];
        rgb[3 * i] = CYCbCr2R(y, u, v);
        rgb[3 * i + 1] = CYCbCr2G(y, u, v);
        rgb[3 * i + 2] = CYCbCr2B(y, u, v);
    }
}
