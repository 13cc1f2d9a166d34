"""f3 on the emulator: osmr_draw_tiles_auto (device-side candidates + painter's order) against osmr_draw_tiles fed with
the host-built lists, on fixture tiles rebuilt from nano_moscow.bin with the committed rules.

    python tests/emu/run_auto.py [zoom] [n_tiles]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)


from autocheck import check_auto, fixture_builder  # noqa: E402


def main():
    from run_emu import emu_context

    zoom = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    from conftest import FixtureInputs

    fx = FixtureInputs()
    tiles = [tuple(int(v) for v in t) for t in fx.batches[str(zoom)][0][:n]]
    data, rd, S, table, fb = fixture_builder()
    ctx = emu_context(os.environ.get("OSMR_EMU_LIB"))
    bad, bad_order, n_dev, n_host = check_auto(ctx, data, S, table, fb, tiles)
    print(f"zoom {zoom}, {len(tiles)} tiles: device lists {n_dev} styled areas (host lists {n_host}), "
          f"order violations {bad_order}, differing pixels auto vs host lists: {bad}")
    return bad + bad_order


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
