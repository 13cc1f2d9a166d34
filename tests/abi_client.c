/* abi_client.c -- a plain C99 client of include/osmr.h (test infrastructure).
 *
 * Compiled by tests/test_host_abi.py with `gcc -std=c99 -Wall -Werror` and linked against libosmr_b200.so: proves that the
 * header is valid C, pins the layout of every struct that crosses the boundary (what a bindgen run over this header would
 * see), and makes one real call sequence through the C ABI.  Without a CUDA device osmr_ctx_create must fail with a
 * negative code and a null context -- never abort (there is no CPU fallback).
 */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "osmr.h"

#define SA(cond, name) typedef char static_assert_##name[(cond) ? 1 : -1]

SA(sizeof(osmr_tile) == 16, tile_size);
SA(offsetof(osmr_tile, zoom) == 0 && offsetof(osmr_tile, x) == 4 && offsetof(osmr_tile, y) == 8 && offsetof(osmr_tile, scale) == 12, tile_fields);

SA(sizeof(osmr_style) == 72, style_size);
SA(offsetof(osmr_style, flags) == 0, style_flags);
SA(offsetof(osmr_style, color) == 4 && offsetof(osmr_style, line_cap) == 7, style_color);
SA(offsetof(osmr_style, fill_color) == 8 && offsetof(osmr_style, casing_line_cap) == 11, style_fill_color);
SA(offsetof(osmr_style, casing_color) == 12 && offsetof(osmr_style, reserved0) == 15, style_casing_color);
SA(offsetof(osmr_style, fill_image) == 16, style_fill_image);
SA(offsetof(osmr_style, width) == 24 && offsetof(osmr_style, opacity) == 32, style_width);
SA(offsetof(osmr_style, fill_opacity) == 40 && offsetof(osmr_style, casing_width) == 48, style_fill_opacity);
SA(offsetof(osmr_style, dashes_off) == 56 && offsetof(osmr_style, dashes_len) == 60, style_dashes);
SA(offsetof(osmr_style, casing_dashes_off) == 64 && offsetof(osmr_style, casing_dashes_len) == 68, style_casing_dashes);

SA(sizeof(osmr_styled_area) == 8 && offsetof(osmr_styled_area, entity) == 0 && offsetof(osmr_styled_area, style) == 4, styled_area);
SA(sizeof(osmr_icon) == 8 + sizeof(void*) && offsetof(osmr_icon, width) == 0 && offsetof(osmr_icon, height) == 4 && offsetof(osmr_icon, rgba) == 8, icon);

SA(sizeof(osmr_label_style) == 32, label_style_size);
SA(offsetof(osmr_label_style, icon) == 0 && offsetof(osmr_label_style, flags) == 4, label_style_head);
SA(offsetof(osmr_label_style, text_key_off) == 8 && offsetof(osmr_label_style, text_key_len) == 12, label_style_key);
SA(offsetof(osmr_label_style, text_color) == 16 && offsetof(osmr_label_style, text_position) == 19, label_style_color);
SA(offsetof(osmr_label_style, reserved0) == 20 && offsetof(osmr_label_style, font_size) == 24, label_style_font);
SA(sizeof(osmr_label) == 8 && offsetof(osmr_label, entity) == 0 && offsetof(osmr_label, style) == 4, label);
SA(sizeof(osmr_class_style) == 8 && offsetof(osmr_class_style, style) == 0 && offsetof(osmr_class_style, order) == 4, class_style);

SA(offsetof(osmr_stats, n_tiles) == 0 && offsetof(osmr_stats, n_areas) == 8 && offsetof(osmr_stats, n_visible_ops) == 16, stats_head);
SA(offsetof(osmr_stats, n_node_refs) == 24 && offsetof(osmr_stats, kernel_launches) == 32 && offsetof(osmr_stats, geom_bytes) == 40, stats_mid);
SA(offsetof(osmr_stats, mask_bytes) == 48 && offsetof(osmr_stats, walk_bytes) == 56 && offsetof(osmr_stats, walk_steps) == 64, stats_walk);
SA(offsetof(osmr_stats, ms_plan) == 72 && offsetof(osmr_stats, ms_raster) == 76 && offsetof(osmr_stats, ms_total) == 80, stats_ms);
SA(offsetof(osmr_stats, ms_label_layout) == 84 && offsetof(osmr_stats, ms_label_device) == 88 && offsetof(osmr_stats, ms_cover) == 92, stats_ms2);
SA(offsetof(osmr_stats, ms_auto) == 96 && offsetof(osmr_stats, ms_png) == 100, stats_ms3);
SA(offsetof(osmr_stats, label_path) == 104 && offsetof(osmr_stats, n_labels_active) == 108, stats_label);
SA(offsetof(osmr_stats, n_labels_polylabel) == 112 && offsetof(osmr_stats, label_attempts) == 116, stats_label2);
SA(offsetof(osmr_stats, ms_label_cover) == 120 && offsetof(osmr_stats, n_label_segments) == 124 && offsetof(osmr_stats, n_label_cells) == 128 && sizeof(osmr_stats) == 136, stats_label3);

int main(void) {
    osmr_ctx* ctx = (osmr_ctx*)1;
    int rc;
    printf("abi_version %u\n", (unsigned)osmr_abi_version());
    printf("png_bound(1) %lu png_bound(0) %lu\n", (unsigned long)osmr_png_bound(1), (unsigned long)osmr_png_bound(0));
    rc = osmr_ctx_create(0, &ctx);
    printf("ctx_create rc %d ctx %s\n", rc, ctx ? "non-null" : "null");
    if (rc != OSMR_OK) {
        if (ctx != NULL) return 2; /* a failed create must null the out pointer */
        if (rc >= 0) return 3;
        /* every entry point tolerates a null context */
        if (osmr_set_geodata(NULL, "", 0) != OSMR_E_INVALID) return 4;
        if (osmr_draw_tiles(NULL, NULL, 0, NULL, NULL, NULL, 0, NULL) != OSMR_E_INVALID) return 5;
        if (strcmp(osmr_last_error(NULL), "null context") != 0) return 6;
        osmr_ctx_destroy(NULL);
        printf("no device: failed loudly\n");
        return 0;
    }
    {
        /* a device is present: the state machine answers, nothing aborts */
        osmr_tile t = {14u, 9903u, 5121u, 1u};
        uint32_t begin[2] = {0u, 0u};
        uint8_t canvas[3] = {241, 238, 232};
        static uint8_t out[256 * 256 * 3];
        rc = osmr_draw_tiles(ctx, &t, 1u, begin, NULL, canvas, OSMR_DRAW_HAS_CANVAS_COLOR, out);
        printf("draw before set_geodata rc %d (%s)\n", rc, osmr_last_error(ctx));
        if (rc != OSMR_E_STATE) return 7;
        osmr_ctx_destroy(ctx);
    }
    return 0;
}
