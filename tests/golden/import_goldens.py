"""Copies the reference's own golden images (test fixtures, not source) into tests/golden/.

Run once in the build container, where the reference checkout is mounted read-only:
    python tests/golden/import_goldens.py /root/reference
The GPU box has no /root/reference, so the fixtures are committed.  Provenance (reference v0.35.1):
    tests/expected/render_rgb_boxes_sdf.png   <- tests/trender_rgb_boxes_sdf.nim:13-101
    tests/expected/render_linear_gradient.png <- tests/trender_linear_gradient.nim:13-96
    tests/expected/render_layers_clip.png     <- tests/trender_layers_clip.nim:76-173
    tests/expected/render_circle_rect.png     <- tests/trender_extras.nim:39-57
    tests/expected/render_line_rect.png       <- tests/trender_extras.nim:18-37
    tests/expected/render_image.png           <- tests/trender_image.nim:13-39 (input: data/img1.png)
`render_rgb_boxes.png` is stale (legacy texture path) and `render_3d_overlay.png` is out of scope.
"""
import os
import shutil
import sys

GOLDENS = ["render_rgb_boxes_sdf.png", "render_linear_gradient.png", "render_layers_clip.png",
           "render_circle_rect.png", "render_line_rect.png", "render_image.png"]

if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    for g in GOLDENS:
        shutil.copyfile(os.path.join(ref, "tests", "expected", g), os.path.join(here, g))
    shutil.copyfile(os.path.join(ref, "data", "img1.png"), os.path.join(here, "img1.png"))
    print("copied", len(GOLDENS) + 1, "fixtures")
