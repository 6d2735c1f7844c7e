/* Plain-C consumer of the drop-in boundary: includes only include/figdraw_cuda.h and links libfigdraw_cuda.so, the way
 * a Nim `importc`, cgo or JNI binding would.  Without a GPU it exercises the host-only entry points (scene flattening)
 * and checks that fdc_create fails cleanly; with a GPU it renders the reference's one-frame-screenshot scene
 * (tests/tfigrender_oneframe_screenshot.nim:20-42) through the per-call API and prints two pixels. */
#include <stdio.h>
#include <string.h>

#include "figdraw_cuda.h"

static fdc_node_fill solid(uint32_t rgba) {
  fdc_node_fill f;
  memset(&f, 0, sizeof(f));
  f.c[0] = rgba;
  f.mid_pos = 128;
  return f;
}

int main(void) {
  printf("abi %d\n", fdc_abi_version());

  /* one rectangle node with a drop shadow and a stroke -> shadow, fill, stroke */
  fdc_fig node;
  memset(&node, 0, sizeof(node));
  node.kind = FDC_NK_RECTANGLE;
  node.parent = -1;
  node.screen_box[0] = 32; node.screen_box[1] = 24; node.screen_box[2] = 120; node.screen_box[3] = 80;
  node.fill = solid(0xFF2828DCu); /* rgba(220, 40, 40, 255) */
  node.u.rect.shadows[0].style = 1;
  node.u.rect.shadows[0].fill = solid(0x5A000000u);
  node.u.rect.shadows[0].blur = 6.0f;
  node.u.rect.shadows[0].spread = 2.0f;
  node.u.rect.stroke.weight = 2.0f;
  node.u.rect.stroke.fill = solid(0xFF000000u);
  int32_t root = 0;
  fdc_render_list list = {&node, 1, &root, 1};
  fdc_scene scene;
  memset(&scene, 0, sizeof(scene));
  scene.lists = &list;
  scene.n_lists = 1;
  fdc_flatten_env env;
  memset(&env, 0, sizeof(env));
  env.ui_scale = 1.0f; env.pixel_scale = 1.0f; env.aa_factor = 1.2f;
  fdc_call calls[16];
  size_t n = 0;
  int rc = fdc_flatten_renders(&scene, &env, calls, 16, &n);
  printf("flatten rc %d records %zu ops", rc, n);
  for (size_t i = 0; i < n; i++) printf(" %u", calls[i].op);
  printf("\n");
  rc = fdc_flatten_renders(&scene, &env, calls, 2, &n);
  printf("flatten small buffer rc %d needed %zu\n", rc, n);

  fdc_ctx* ctx = NULL;
  rc = fdc_create(&ctx, 0, 512, 1.0f, 0, 1);
  if (rc != FDC_OK) {
    printf("create rc %d (no usable GPU): %s\n", rc, fdc_last_error(NULL));
    return 0;
  }
  const float white[4] = {1, 1, 1, 1};
  const float zero4[4] = {0, 0, 0, 0};
  const float full[4] = {0, 0, 240, 160}, box[4] = {32, 24, 120, 80};
  fdc_fill bg, red;
  memset(&bg, 0, sizeof(bg));
  memset(&red, 0, sizeof(red));
  bg.kind = FDC_FILL_COLOR; bg.c[0] = 0xFFFFFFFFu; bg.mid_pos = 0.5f;
  red.kind = FDC_FILL_COLOR; red.c[0] = 0xFF2828DCu; red.mid_pos = 0.5f;
  const float no_shape[2] = {0, 0};
  rc = fdc_begin_frame(ctx, 240, 160, 1, white);
  if (!rc) rc = fdc_draw_rounded_rect_sdf(ctx, full, &bg, zero4, zero4, FDC_SDF_CLIP_AA, 4.0f, 0.0f, no_shape);
  if (!rc) rc = fdc_draw_rounded_rect_sdf(ctx, box, &red, zero4, zero4, FDC_SDF_CLIP_AA, 4.0f, 0.0f, no_shape);
  if (!rc) rc = fdc_end_frame(ctx);
  static uint8_t px[160 * 240 * 4];
  if (!rc) rc = fdc_read_pixels(ctx, 0, 0, 240, 160, px);
  if (rc) {
    printf("render failed rc %d: %s\n", rc, fdc_last_error(ctx));
    fdc_destroy(ctx);
    return 1;
  }
  const uint8_t* a = px + (12 * 240 + 12) * 4;
  const uint8_t* b = px + (48 * 240 + 64) * 4;
  printf("pixel(12,12) %u %u %u %u pixel(64,48) %u %u %u %u\n", a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]);
  fdc_destroy(ctx);
  return 0;
}
