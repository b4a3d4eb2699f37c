/*
 * ref_glue.cpp -- C entry points onto the REFERENCE's own classes, compiled together with the reference
 * sources where they lie under /root/reference (never copied).  TEST INFRASTRUCTURE ONLY.
 * Output: oracle/_ref/libjade_ref.so (git-ignored; travels to the GPU box as a built file).
 *
 *   jr_pal_*  -> CColorPalette            (CColorpalette.h:20-48, CColorpalette.cpp)
 */
#include "CColorpalette.h"

extern "C" {
void* jr_pal_create(int n, int scheme) { return new CColorPalette(n, scheme); }
void* jr_pal_create_default(void) { return new CColorPalette(); }
void jr_pal_destroy(void* p) { delete static_cast<CColorPalette*>(p); }
void jr_pal_set_value_range(void* p, float a, float b) { static_cast<CColorPalette*>(p)->setValueRange(a, b); }
void jr_pal_set_nr_of_colors(void* p, int n) { static_cast<CColorPalette*>(p)->setNrOfColors(n); }
void jr_pal_set_color_scheme(void* p, int s) { static_cast<CColorPalette*>(p)->setColorSceme(s); }
void jr_pal_set_invert(void* p, int on) { static_cast<CColorPalette*>(p)->setInvertStatus(on != 0); }
int jr_pal_get_rgb(void* p, float v) { return static_cast<CColorPalette*>(p)->getRGBColor(v); }
float jr_pal_get_value(void* p, int c) { return static_cast<CColorPalette*>(p)->getValue(c); }
void jr_pal_lookup_many(void* p, const float* v, int n, int* out)
{
    CColorPalette* c = static_cast<CColorPalette*>(p);
    for (int i = 0; i < n; ++i) out[i] = c->getRGBColor(v[i]);
}
}
