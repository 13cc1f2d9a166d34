/*
 * osmr.h -- C ABI of the B200-native tile rasteriser (libosmr_b200.so).
 *
 * Drop-in boundary for ONE path of dfyz/osm-renderer: everything between "ordered styled areas" and
 * "RGB triples" inside `Drawer::draw_to_pixels` (reference src/draw/drawer.rs:60-131), i.e.
 *   a1 Web-Mercator lon/lat -> integer tile pixel ... src/tile.rs:88-106, src/draw/point.rs:11-19
 *   a2/a3 even-odd polygon scanline fill ............. src/draw/fill.rs:16-104
 *   a4/a5 thick anti-aliased lines, dashes, caps ...... src/draw/line.rs:9-158, opacity_calculator.rs:16-185
 *   a6/a7 generation compositor, alpha-over, export .... src/draw/tile_pixels.rs:89-129,164-181,205-223
 *   a8 Fill -> Casing -> Stroke pass ordering ........... src/draw/drawer.rs:94-100,133-219
 * The styler (string matching, reference src/mapcss/styler.rs) stays on the host and hands over its
 * output unchanged: an ordered list of (area, style) pairs per tile.  Batching many tiles per call is the
 * only semantic extension; n_tiles == 1 reproduces one reference call.
 *
 * Conventions: every function returns 0 on success or a negative OSMR_E_* code and never aborts or throws
 * across the boundary; the message of the last failure on a context is osmr_last_error(ctx).  A context is
 * NOT thread-safe (one per host worker thread, the analogue of the reference's per-thread TilePixels,
 * src/http_server.rs:69-72).  No pointer argument is retained after a call returns (data is copied to the
 * device).  There is no CPU fallback: without a usable CUDA device osmr_ctx_create fails.
 */
#ifndef OSMR_H
#define OSMR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSMR_OK 0
#define OSMR_E_INVALID (-1)  /* bad argument / malformed geodata image */
#define OSMR_E_CUDA (-2)     /* CUDA runtime error (see osmr_last_error) */
#define OSMR_E_NOMEM (-3)
#define OSMR_E_STATE (-4)    /* call sequence error (e.g. draw before osmr_set_geodata) */

/* reference: struct Tile {zoom: u8, x: u32, y: u32} (src/tile.rs:8-13) + the `scale: usize` argument of
 * Drawer::draw_to_pixels (src/draw/drawer.rs:65).  Output dimension D = 256 * scale. */
typedef struct osmr_tile {
    uint32_t zoom;
    uint32_t x;
    uint32_t y;
    uint32_t scale;
} osmr_tile;

/* Option<...> presence bits of osmr_style.flags */
#define OSMR_STYLE_COLOR (1u << 0)
#define OSMR_STYLE_FILL_COLOR (1u << 1)
#define OSMR_STYLE_FILL_IMAGE (1u << 2)
#define OSMR_STYLE_CASING_COLOR (1u << 3)
#define OSMR_STYLE_CASING_WIDTH (1u << 4)
#define OSMR_STYLE_WIDTH (1u << 5)
#define OSMR_STYLE_OPACITY (1u << 6)
#define OSMR_STYLE_FILL_OPACITY (1u << 7)
#define OSMR_STYLE_DASHES (1u << 8)
#define OSMR_STYLE_CASING_DASHES (1u << 9)

/* reference: enum LineCap (src/mapcss/styler.rs:11-16); 0 encodes Option::None */
#define OSMR_CAP_NONE 0
#define OSMR_CAP_BUTT 1
#define OSMR_CAP_ROUND 2
#define OSMR_CAP_SQUARE 3

/* The draw-relevant fields of reference `struct Style` (src/mapcss/styler.rs:49-72).  Widths and dashes are
 * UNSCALED (the library multiplies by tile.scale exactly as drawer.rs:163-164,193,207 does).
 * fill_image is an index into the icon table of osmr_set_icons, or -1 when the icon failed to load
 * (the reference then silently skips the area: drawer.rs:179-184, icon_cache.rs:32-41). */
typedef struct osmr_style {
    uint32_t flags;
    uint8_t color[3];
    uint8_t line_cap;
    uint8_t fill_color[3];
    uint8_t casing_line_cap;
    uint8_t casing_color[3];
    uint8_t reserved0;
    int32_t fill_image;
    double width;
    double opacity;
    double fill_opacity;
    double casing_width;
    uint32_t dashes_off; /* into the `dashes` array of osmr_set_styles */
    uint32_t dashes_len;
    uint32_t casing_dashes_off;
    uint32_t casing_dashes_len;
} osmr_style;

/* One element of the styler's output Vec<(StyledArea, Arc<Style>)> (src/mapcss/styler.rs:168-203).
 * entity: local index into the `.bin` way table, or OSMR_AREA_MULTIPOLYGON | index into the multipolygon
 * table (reference enum StyledArea, styler.rs:85-92). */
#define OSMR_AREA_MULTIPOLYGON 0x80000000u
typedef struct osmr_styled_area {
    uint32_t entity;
    uint32_t style;
} osmr_styled_area;

/* reference: struct Icon (src/draw/icon.rs:8-12) before premultiplication: 8-bit RGBA rows, as produced by
 * png `normalize_to_color8` + the colour-type switch of icon.rs:29-49. */
typedef struct osmr_icon {
    uint32_t width;
    uint32_t height;
    const uint8_t* rgba; /* width*height*4 bytes */
} osmr_icon;

/* draw flags */
#define OSMR_DRAW_USE_CAPS_FOR_DASHES (1u << 0) /* Styler.use_caps_for_dashes (styler.rs:95) */
#define OSMR_DRAW_HAS_CANVAS_COLOR (1u << 1)    /* Styler.canvas_fill_color is Some (styler.rs:96) */
#define OSMR_DRAW_OUT_RGBA (1u << 2)            /* 4 bytes/pixel, A=255; default is the reference's RGB triples */
#define OSMR_DRAW_OUT_DEVICE (1u << 3)          /* `out` is a device pointer (stays in HBM) */

typedef struct osmr_ctx osmr_ctx;

/* replaces Drawer::new + TilePixels::new (drawer.rs:33, tile_pixels.rs:57): scratch + CUDA stream on `device` */
int osmr_ctx_create(int device, osmr_ctx** out_ctx);
/* A worker context that SHARES the parent's resident dataset (reference: src/http_server.rs:42-48,69-72 -- GeodataReader, Styler
 * and Drawer are shared immutably by all worker threads, only TilePixels is per thread).  Own streams, scratch, style / icon /
 * label tables and zoom classes; the geodata uploaded by osmr_set_geodata on the parent is used by reference (no second copy in
 * HBM, no second upload).  The dataset is immutable: osmr_set_geodata on either context gives that context a dataset of its own.
 * Each context is still used by one host thread at a time; different contexts may run on different threads concurrently. */
int osmr_ctx_create_shared(const osmr_ctx* parent, osmr_ctx** out_ctx);
void osmr_ctx_destroy(osmr_ctx* ctx);
const char* osmr_last_error(const osmr_ctx* ctx);

/* replaces GeodataReader::load (src/geodata/reader.rs:44-58): `bin` is the geodata file image in the
 * reference's on-disk format (saver.rs:21-165).  The zoom-independent Mercator factor of every node (a1, transcendental
 * part) is computed once -- with the platform libm, like the reference -- and kept resident together with the way / polygon /
 * multipolygon tables, the entities' bounding boxes and the tile index. */
int osmr_set_geodata(osmr_ctx* ctx, const void* bin, size_t bin_len);

/* replaces IconCache for `fill-image` patterns (src/draw/icon_cache.rs:21-45, icon.rs:14-58) */
int osmr_set_icons(osmr_ctx* ctx, const osmr_icon* icons, uint32_t n_icons);

/* the interned Arc<Style> table (styler.rs:49-72) + flat dash storage */
int osmr_set_styles(osmr_ctx* ctx, const osmr_style* styles, uint32_t n_styles, const double* dashes, uint32_t n_dashes);

/* replaces Drawer::draw_to_pixels for the area passes (drawer.rs:60-104,128): for tile t the styled areas are
 * areas[area_begin[t] .. area_begin[t+1]) in styler order.  canvas_rgb = Styler.canvas_fill_color when
 * OSMR_DRAW_HAS_CANVAS_COLOR is set.  `out` receives n_tiles images of D*D*3 (or *4) bytes, tightly packed in
 * tile order, D = 256*scale; all tiles of one call must share one scale. Host pointers unless OUT_DEVICE. */
int osmr_draw_tiles(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                    const osmr_styled_area* areas, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out);

/* Same work with the batch description already resident in HBM (benchmark "value" leg): upload once ... */
int osmr_batch_upload(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                      const osmr_styled_area* areas);
/* ... then draw any number of times; `out` as in osmr_draw_tiles (NULL = keep the result in the context's
 * own device buffer).  gpu_ms (optional) receives the device time of this call measured with CUDA events
 * on the context's stream. */
int osmr_batch_draw(osmr_ctx* ctx, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out, float* gpu_ms);

/* device pointer + byte size of the context-owned output of the last draw (RGB or RGBA as drawn) */
int osmr_batch_output(osmr_ctx* ctx, const uint8_t** dev_ptr, size_t* n_bytes);

/* counters of the last draw, for the roofline arithmetic of bench.py */
typedef struct osmr_stats {
    uint64_t n_tiles;
    uint64_t n_areas;        /* styled areas in the batch (G/3 generations per pass) */
    uint64_t n_visible_ops;  /* generations whose reach intersects their tile */
    uint64_t n_node_refs;    /* R: node references of all styled areas */
    uint64_t kernel_launches;
    uint64_t geom_bytes;     /* edge/segment records written for the visible ops */
    uint64_t mask_bytes;     /* fill row masks written (1 bit per pixel of every fill row) */
    uint64_t walk_bytes;     /* walk cache handed out (8 bytes per possible step of every perpendicular walk) */
    uint64_t walk_steps;     /* in-line walk steps actually evaluated and stored by line_cover_kernel */
    float ms_plan;           /* per-stage device times of the last draw (CUDA events): bbox, plan, geometry, fill rows */
    float ms_raster;
    float ms_total;
    float ms_label_layout;   /* osmr_draw_tiles_labeled: host layout (wall clock) and label kernels (CUDA events) */
    float ms_label_device;
    float ms_cover;          /* line_cover_kernel */
    float ms_auto;           /* osmr_draw_tiles_auto: candidate lookup + ordering on the device */
    float ms_png;            /* osmr_draw_tiles_png: filter + deflate + checksums on the device */
    uint32_t label_path;     /* osmr_draw_tiles_labeled: 0 no label pass, 1 layout on the device, 2 layout on the host (fallback / debug key) */
    uint32_t n_labels_active;    /* device layout: label generations with an icon or an existing text */
    uint32_t n_labels_polylabel; /* device layout: of those, areas whose anchor is the pole of inaccessibility (labelable.rs:125-189) */
    uint32_t label_attempts;     /* device layout: attempts of this call (> 1: scratch was grown and the call redone) */
    float ms_label_cover;        /* device layout: label_cover_kernel alone (the longest kernel of a labelled draw) */
    uint32_t n_label_segments;   /* device layout: glyph outline segments (draw_line calls of the reference's rasteriser) */
    uint64_t n_label_cells;      /* device layout: coverage cells of the labels that touch the label canvas */
} osmr_stats;
int osmr_get_stats(osmr_ctx* ctx, osmr_stats* out);

/* a1 in isolation (parity tests): integer pixel of every node for one tile, exactly
 * Point::from_node (point.rs:11-19).  out_xy: n_nodes * 2 int32 (host). */
int osmr_project_nodes(osmr_ctx* ctx, const osmr_tile* tile, int32_t* out_xy);

/* ------------------------------------------------------------------------------------------------------------
 * Label pass (reference src/draw/drawer.rs:106-126,221-262; src/draw/labeler.rs:16-106; src/draw/labelable.rs;
 * src/draw/font/{text_placer,rasterizer}.rs; src/draw/tile_pixels.rs:131-162,205-209).
 * String / font-table / libm work is done by host C++ inside the library ONCE, as resident tables: per (dataset, font, label
 * style table) the text run of every entity (tag lookup, cmap, hmtx, kern) and the stb_truetype glyph outlines; per zoom the
 * sin / cos of every named way's segment directions (glibc, like the reference).  Everything per tile and per call runs on
 * the device: which generations can draw, polylabel anchors, glyph placement along ways / in wrapped rows, outline flattening,
 * exact-area glyph coverage in segment order, the greedy label collisions over the 3x3 canvas, and the blend of the surviving
 * labels over the f64 tile canvas before the RGB export.  Inputs the device path does not take (a scale that is not a power of
 * two, zoom > 18, a flatness near-tie that only libm's hypot can decide, ...) are laid out by the host code of round 1 --
 * same pixels either way; osmr_stats.label_path says which ran.
 * ------------------------------------------------------------------------------------------------------------ */
#define OSMR_LABEL_NODE 0x40000000u /* osmr_label.entity: node index | OSMR_LABEL_NODE (ways / multipolygons as in
                                       osmr_styled_area) */
#define OSMR_LSTYLE_TEXT (1u << 0)       /* Style.text_style is Some */
#define OSMR_LSTYLE_FONT_SIZE (1u << 1)  /* TextStyle.font_size is Some (already multiplied by font-mul) */
#define OSMR_LSTYLE_TEXT_COLOR (1u << 2) /* TextStyle.text_color is Some (default black, text_placer.rs:52-55) */
#define OSMR_TEXT_POS_NONE 0
#define OSMR_TEXT_POS_CENTER 1
#define OSMR_TEXT_POS_LINE 2

/* label-relevant fields of reference `struct Style` / `struct TextStyle` (styler.rs:42-47,69-71) */
typedef struct osmr_label_style {
    int32_t icon;          /* -2: no icon-image; -1: icon failed to load (labeler.rs:52-67 then continues with the
                              text); >= 0: index into the table of osmr_set_label_icons */
    uint32_t flags;        /* OSMR_LSTYLE_* */
    uint32_t text_key_off; /* TextStyle.text = the tag KEY whose value is drawn; bytes in `strings` */
    uint32_t text_key_len;
    uint8_t text_color[3];
    uint8_t text_position; /* OSMR_TEXT_POS_* (NONE = default: Line for ways, Center otherwise, drawer.rs:232-257) */
    uint32_t reserved0;
    double font_size;
} osmr_label_style;

/* one label generation: element of the styler output for labels, in the reference's order (areas styled with
 * for_labels = true, then nodes; drawer.rs:106-119) */
typedef struct osmr_label {
    uint32_t entity;
    uint32_t style; /* index into the table of osmr_set_label_styles */
} osmr_label;

/* replaces `FONT_DATA` (text_placer.rs:299): the TrueType file used for every label */
int osmr_set_font(osmr_ctx* ctx, const void* ttf, size_t len);
/* replaces IconCache for `icon-image` (labeler.rs:46-50); same pixel format as osmr_set_icons */
int osmr_set_label_icons(osmr_ctx* ctx, const osmr_icon* icons, uint32_t n_icons);
int osmr_set_label_styles(osmr_ctx* ctx, const osmr_label_style* styles, uint32_t n_styles, const char* strings, size_t strings_len);
/* osmr_draw_tiles including the label pass: the labels of tile t are labels[label_begin[t] .. label_begin[t+1]).
 * Requires osmr_set_font (and the label tables) first. */
int osmr_draw_tiles_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                            const osmr_styled_area* areas, const uint32_t* label_begin, const osmr_label* labels,
                            const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out);

/* The same call with everything resident in HBM (benchmark "value" leg of the whole draw_to_pixels): upload once, draw any number
 * of times; `out` / `gpu_ms` as in osmr_batch_draw.  Needs the device label layout (scale 1, 2, 4 or 8; zoom <= 18). */
int osmr_batch_upload_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                              const osmr_styled_area* areas, const uint32_t* label_begin, const osmr_label* labels);
int osmr_batch_draw_labeled(osmr_ctx* ctx, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out, float* gpu_ms);

/* ------------------------------------------------------------------------------------------------------------
 * the step BEFORE the draw path on the device -- tile -> candidate entities -> ordered styled areas.
 * Replaces GeodataReader::get_entities_in_tile_with_neighbors (src/geodata/reader.rs:60-180), Styler::style_areas
 * (src/mapcss/styler.rs:115-203) and the painter's order of compare_styled_entities (styler.rs:246-272) for the area
 * passes.  MapCSS selector matching (strings) stays on the host, exactly as cached by the reference's StyleCache
 * (src/mapcss/style_cache.rs:68-87, key = entity + zoom): per zoom the host provides, for every way and multipolygon of
 * the `.bin`, the id of its style list ("class", 0xffffffff = no styles) and per class the ordered list of its styles.
 * osmr_class_style.order is the dense rank (equal keys -> equal rank) of the style's sort key
 * (layer.unwrap_or(0), is_foreground_fill, z_index) among all class styles of the zoom; the library appends the rest of
 * the reference's comparison (global id, multipolygon before way, local id, position in the style list) itself.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct osmr_class_style {
    uint32_t style; /* index into the table of osmr_set_styles */
    uint32_t order; /* < 2^20 */
} osmr_class_style;
int osmr_set_zoom_styles(osmr_ctx* ctx, uint32_t zoom, const uint32_t* way_class /* n_ways */, const uint32_t* mp_class /* n_multipolygons */,
                         const uint32_t* class_begin /* n_classes + 1 */, const osmr_class_style* class_styles, uint32_t n_classes);
/* osmr_draw_tiles with the styled-area lists built on the device: only the tile list crosses the bus.  All tiles of a call
 * share one zoom (<= 18, tile.rs:5) and one scale.  Areas whose bounding box (plus the widest line reach of their styles)
 * cannot touch the tile are dropped before ordering; the image is the one osmr_draw_tiles produces from the reference's
 * full list. */
int osmr_draw_tiles_auto(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out);
/* the styled-area lists of the last osmr_draw_tiles_auto call (tests): area_begin[n_tiles + 1], areas[area_begin[n_tiles]] */
int osmr_auto_readback(osmr_ctx* ctx, uint32_t* area_begin, osmr_styled_area* areas, uint32_t areas_cap);

/* f3 for the LABEL pass: what the reference does between the area passes and draw_labels (src/draw/drawer.rs:106-119) --
 * Styler::style_areas(.., for_labels = true) (src/mapcss/styler.rs:168-203) and Styler::style_entities(nodes) (styler.rs:115-166)
 * over the entities of GeodataReader::get_entities_in_tile_with_neighbors (src/geodata/reader.rs:60-100, nodes included) -- on the
 * device.  As for the area passes the host provides the StyleCache contents per zoom: for every node, way and multipolygon of
 * the `.bin` the id of its style list (0xffffffff = none), per class its styles as (index into the table of
 * osmr_set_label_styles, order) with `order` (< 2^19) the dense rank of (layer.unwrap_or(0), z_index) -- compare_styled_entities
 * ignores is_foreground_fill when for_labels is set (styler.rs:263).  The library appends (global id, multipolygon before way,
 * local id, position in the style list) and lists the styled nodes behind the styled areas. */
int osmr_set_zoom_label_styles(osmr_ctx* ctx, uint32_t zoom, const uint32_t* node_class /* n_nodes */, const uint32_t* way_class /* n_ways */,
                               const uint32_t* mp_class /* n_multipolygons */, const uint32_t* class_begin /* n_classes + 1 */,
                               const osmr_class_style* class_styles, uint32_t n_classes);
/* The whole Drawer::draw_to_pixels (src/draw/drawer.rs:60-131) from a tile list: styled-area lists AND label lists are built on
 * the device, then osmr_draw_tiles_labeled's kernels run on them.  Needs osmr_set_zoom_styles, osmr_set_zoom_label_styles,
 * osmr_set_font and the label tables; conditions as osmr_draw_tiles_auto.  Label generations that cannot draw or collide (no
 * icon, no text for the entity) are not listed -- the reference draws nothing for them (src/draw/labeler.rs:16-37). */
int osmr_draw_tiles_auto_labeled(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* out);
/* the label lists of the last osmr_draw_tiles_auto_labeled* call (tests): label_begin[n_tiles + 1], labels[label_begin[n_tiles]] */
int osmr_auto_readback_labels(osmr_ctx* ctx, uint32_t* label_begin, osmr_label* labels, uint32_t labels_cap);

/* ------------------------------------------------------------------------------------------------------------
 * the step AFTER the draw path -- PNG files instead of RGB triples.
 * Replaces Drawer::draw_tile = draw_to_pixels + rgb_triples_to_png (src/draw/drawer.rs:40-58, src/draw/png_writer.rs:4-21:
 * 8-bit RGB, one IDAT chunk).  Filter choice and deflate stream are the encoder's (per-row None/Sub/Up/Paeth by smallest
 * absolute residual sum, fixed Huffman code, run-length matches), the decoded image is exactly the RGB tile.
 * The files of the batch are packed back to back into png_out: tile t occupies [png_offset[t], png_offset[t+1]).
 * png_cap >= n_tiles * osmr_png_bound(scale) always suffices (the call fails with OSMR_E_NOMEM otherwise, png_offset
 * then still holds the sizes).
 * ------------------------------------------------------------------------------------------------------------ */
size_t osmr_png_bound(uint32_t scale);
int osmr_draw_tiles_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* area_begin,
                        const osmr_styled_area* areas, const uint8_t canvas_rgb[3], uint32_t flags, uint8_t* png_out, size_t png_cap,
                        uint64_t* png_offset /* n_tiles + 1 */);

/* f3 + f4 around the draw path = Drawer::draw_tile as the server calls it (src/http_server.rs:150-177: entities of the tile,
 * draw_tile, PNG bytes to the client): only the tile list goes to the device, only PNG files come back.  Same conditions as
 * osmr_draw_tiles_auto (one zoom and one scale per call, osmr_set_zoom_styles first), output as osmr_draw_tiles_png. */
int osmr_draw_tiles_auto_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags,
                             uint8_t* png_out, size_t png_cap, uint64_t* png_offset /* n_tiles + 1 */);
/* the same with the label pass (osmr_draw_tiles_auto_labeled in front of the encoder): Drawer::draw_tile complete */
int osmr_draw_tiles_auto_labeled_png(osmr_ctx* ctx, const osmr_tile* tiles, uint32_t n_tiles, const uint8_t canvas_rgb[3], uint32_t flags,
                                     uint8_t* png_out, size_t png_cap, uint64_t* png_offset /* n_tiles + 1 */);

/* rgb_triples_to_png itself (png_writer.rs:4-21) for images the caller already holds: n_images RGB images of
 * (256 * scale)^2 pixels, tightly packed in host memory -> packed PNG files as in osmr_draw_tiles_png. */
int osmr_rgb_to_png(osmr_ctx* ctx, const uint8_t* rgb, uint32_t n_images, uint32_t scale, uint8_t* png_out, size_t png_cap,
                    uint64_t* png_offset /* n_images + 1 */);

/* Optional page-locked host memory for callers that want full-speed host<->device copies of `out` / batch
 * arrays (plain malloc'ed buffers work too, at pageable-copy speed). */
void* osmr_alloc_pinned(size_t bytes);
void osmr_free_pinned(void* p);

/* Test / measurement hooks (never needed for correct results):
 *   "fill_cap" (0..128)     lowers the number of row spans ranked in shared memory so that the order-free streaming form
 *                           of the even-odd rule is exercised on ordinary data (also switches the row-parallel path off);
 *   "scratch_units" (n)     restarts the bump-allocated geometry / mask / walk-cache scratch at n units and "work_items" (n)
 *                           pretends the per-op work lists hold n items, so the grow-and-redo paths run;
 *   "plan_slice_areas" (n)  styled areas per slice of plan_ops_kernel's (tile, pass) lists (0: default 16384; small values
 *                           exercise the sliced form on ordinary tiles);
 *   "label_cull" (0/1)      labels that cannot reach the tile get no layout, outlines or coverage (default 1; same pixels);
 *   "label_serial" (0/1)    a resident labelled draw finishes its label pass before the area passes start (timing of the label
 *                           kernels with nothing beside them; default 0);
 *   "label_threads" (1..256) host threads of the label layout;
 *   "host_chunks" (0..16)   equal draw chunks of a host-output call (0: the tapered default schedule);
 *   "two_streams" (0/1)     draw chunks alternate between two compute streams (default 1);
 *   "resident_chunks" (1..16) draw chunks when the output stays in HBM (default 1);
 *   "direct_out" (0/1)      raster_kernel stores the tiles straight into a page-locked `out` (default 0: staged D2H);
 *   "device_merc" (0/1)     per-node Mercator factors by the device's tan / log at the next osmr_set_geodata (default 0: the
 *                           host libm, i.e. the reference's own values to the last bit). */
int osmr_debug_set(osmr_ctx* ctx, const char* key, int value);

uint32_t osmr_abi_version(void); /* 5: osmr_set_zoom_label_styles, osmr_draw_tiles_auto_labeled{,_png}, osmr_auto_readback_labels (4: osmr_stats gained the
                                    label_* fields; osmr_draw_tiles_auto_png, osmr_ctx_create_shared, osmr_batch_*_labeled) */

#ifdef __cplusplus
}
#endif
#endif /* OSMR_H */
