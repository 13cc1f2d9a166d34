// Dev tool (CPU only): times osmr_host::layout_tile on the fixture label lists.  Built by tools/dev/layout_prof.py.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "osmr.h"
#include "osmr_labels_host.hpp"

extern "C" double layout_prof(const uint8_t* bin, size_t bin_len, const uint8_t* ttf, size_t ttf_len, const osmr_label_style* styles,
                              uint32_t n_styles, const char* strings, const uint32_t* icon_wh, uint32_t n_icons,
                              const osmr_tile* tiles, uint32_t n_tiles, const uint32_t* label_begin, const osmr_label* labels,
                              int reps, uint64_t* out_counts) {
    osmr_host::BinView view;
    if (!view.parse(bin, bin_len)) return -1;
    osmr_host::TrueType font;
    if (!font.load(ttf, ttf_len)) return -2;
    std::vector<osmr_host::LabelStyleHost> st(n_styles);
    for (uint32_t i = 0; i < n_styles; ++i) {
        st[i].s = styles[i];
        if (styles[i].flags & OSMR_LSTYLE_TEXT) st[i].key.assign(strings + styles[i].text_key_off, styles[i].text_key_len);
    }
    std::vector<osmr_host::IconDim> dims(n_icons);
    for (uint32_t i = 0; i < n_icons; ++i) dims[i] = osmr_host::IconDim{icon_wh[2 * i], icon_wh[2 * i + 1]};
    osmr_host::LayoutEnv env{&view, &font, &st, &dims};
    auto t0 = std::chrono::steady_clock::now();
    size_t nrec = 0, nseg = 0;
    for (int r = 0; r < reps; ++r)
        for (uint32_t t = 0; t < n_tiles; ++t) {
            std::vector<osmr_host::LabelRec> recs;
            std::vector<osmr_host::Seg> segs;
            if (!osmr_host::layout_tile(env, tiles[t], labels + label_begin[t], label_begin[t + 1] - label_begin[t], recs, segs)) return -3;
            nrec += recs.size();
            nseg += segs.size();
        }
    out_counts[0] = nrec / reps;
    out_counts[1] = nseg / reps;
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
}
