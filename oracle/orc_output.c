/*
 * Output path after the DPB: dynamic-range adjustment on pull (Main, tool_dra), cropping, 16 -> 8-bit conversion.
 * TEST INFRASTRUCTURE ONLY (orc_common.h).  Restates xevd_apply_filter -> xevd_apply_dra_chroma_plane / _luma_plane
 * (src_main/xevdm.c:3311-3348, src_main/xevdm_dra.c:272-354: chroma first, scaled around 512 by a factor looked up from the co-located
 * UNMAPPED luma sample, then luma through its inverse LUT), the crop window xevd_pull_frm attaches (src_main/xevdm.c:3366-3373) and
 * the application's imgb_conv_16b_to_8b (app/xevd_app_util.h:359-381: (v + 2) >> 2, clipped to 0..255).
 */
#include <string.h>
#include "orc_common.h"

void orc_dra_apply(ORC_PIC *pic, const XB200_DRA *d)
{
    for (int c = 0; c < 2; c++) {
        pel *pl = c ? pic->v : pic->u;
        for (int j = 0; j < pic->h_c; j++)
            for (int k = 0; k < pic->w_c; k++) {
                int ref = pic->y[(2 * j) * pic->s_l + 2 * k];
                if (ref < 0) ref = 0;
                const int16_t sv = (int16_t)(pl[j * pic->s_c + k] - 512);
                int off = sv < 0 ? -sv : sv;
                off = (off * d->chroma_inv_scale_lut[c][ref] + 256) >> 9;
                pl[j * pic->s_c + k] = (pel)(512 + (sv < 0 ? -off : off));
            }
    }
    for (int j = 0; j < pic->h_l; j++)
        for (int k = 0; k < pic->w_l; k++) pic->y[j * pic->s_l + k] = (pel)d->luma_inv_scale_lut[pic->y[j * pic->s_l + k]];
}

/* cropped planes, 16-bit (out_bits == 16) or 8-bit; strides in samples */
void orc_output(const ORC_PIC *pic, int out_bits, int crop_l, int crop_r, int crop_t, int crop_b, void *y, int sy, void *u, int su, void *v, int sv)
{
    for (int c = 0; c < 3; c++) {
        const int sh = c ? 1 : 0;
        const pel *src = c == 0 ? pic->y : (c == 1 ? pic->u : pic->v);
        const int s = c ? pic->s_c : pic->s_l;
        const int x0 = crop_l >> sh, y0 = crop_t >> sh, w = ((pic->w_l - crop_l - crop_r) >> sh), h = ((pic->h_l - crop_t - crop_b) >> sh);
        void *dst = c == 0 ? y : (c == 1 ? u : v);
        const int ds = c == 0 ? sy : (c == 1 ? su : sv);
        for (int j = 0; j < h; j++)
            for (int k = 0; k < w; k++) {
                const int val = src[(y0 + j) * s + x0 + k];
                if (out_bits == 16) ((int16_t *)dst)[j * ds + k] = (int16_t)val;
                else ((uint8_t *)dst)[j * ds + k] = (uint8_t)orc_clip3(0, 255, (val + 2) >> 2);
            }
    }
}
