/*
 * oracle/raster_ref.c -- CPU restatement of the renderer contract.  TEST INFRASTRUCTURE ONLY:
 * nothing under megapose6d_b200/ may import, link or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as
 * the checker or the timed CPU baseline.
 *
 * What it restates: the reference renders through Panda3D/OpenGL
 *   src/megapose/panda3d_renderer/panda3d_batch_renderer.py:89-150,217-282 (one view per
 *     (label, TCO, K); non-finite pose or K => all-zero view)
 *   src/megapose/panda3d_renderer/panda3d_scene_renderer.py:298-358 (pass 1: albedo under a single
 *     ambient light of 1.0; pass 2: eye normal looked up in a 32^3 texture), :99-101 two-sided
 *   src/megapose/panda3d_renderer/types.py:58-101 (pinhole lens from K, near 0.1, far 10)
 *   src/megapose/panda3d_renderer/utils.py:44-68 (depth = a/(d-b), d > 0.999 => 0; texel
 *     (x,y,z) = uint8((x,y,z)*255/32))
 * Panda3D is a third-party engine that is neither in /root/reference nor installable here
 * (conda/environment_full.yaml:40 `panda3d`, unpinned; docker builds github.com/ylabbe/panda3d@rebase),
 * and the reference ships no rendering test or golden image.  PARITY UNPINNED: this file defines the
 * geometric contract that the CUDA rasteriser implements bit for bit:
 *   - pixel (i, j) sampled at (u, v) = (j + 0.5, i + 0.5), u = fx X/Z + cx, v = fy Y/Z + cy
 *   - vertices snapped to 1/256 pixel, exact integer edge functions, inclusive edges, two-sided
 *   - barycentrics l_k = w_k * (1/area); the fragment with the largest interpolated 1/z wins, ties -> lower
 *     triangle index; fragments with 1/z outside [1/10, 1/0.1] rejected -- 1/z is linear in screen space, so this per-sample
 *     test IS the clip against the near (z = 0.1) and far (z = 10) planes of the reference's lens (types.py:63-64, :77-80):
 *     a triangle that straddles the near plane keeps exactly its part beyond it.  Only a triangle with a vertex within
 *     2^-10 m (1 mm) of the eye plane, or behind it, has no projection and is dropped (the eye inside the surface)
 *   - 1/z linear in screen space, attributes perspective-correct (b_k = l_k/z_k * z, z = 1/(1/z))
 *   - single sample per pixel (the reference's 4x MSAA is not modelled)
 *   - textured meshes (panda3d_scene_renderer.py:195-208 loads the model with its material / texture): per-vertex
 *     (u, v) interpolated like the other attributes, wrapped to [0, 1) (repeat), v up (image row = (1 - v) * th - 0.5),
 *     bilinear filter over texel centres without mip-mapping, texel = byte / 255; the texture replaces the albedo, or
 *     modulates the interpolated vertex colours when the mesh has them
 *   - flags bit 2 (4): the albedo is lit by make_scene_lights() (panda3d_scene_renderer.py:104-136: ambient 0.1 + six white
 *     point lights of 0.4 at +-10 r on the object's axes, r = max |vertex|) instead of ambient 1.0: Lambert term per
 *     fragment from the perspective-correct position and the normalised interpolated normal in the object frame, no
 *     attenuation, no normal flip on back faces: rgb = albedo * (0.1 + 0.4 * sum_i max(0, n . (L_i - p) / |L_i - p|))
 * Every float operation is a single correctly-rounded IEEE operation in a fixed order (compile with
 * -ffp-contract=off) so that the device kernel can reproduce the results exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define K_NEAR 0.1f
#define K_PROJ_MIN 0.0009765625f /* 2^-10 m: vertices nearer to the eye plane than this (or behind it) cannot be projected */
#define K_IZ_MAX 10.0f /* 1 / near */
#define K_IZ_MIN 0.1f  /* 1 / far */
#define K_SUB 256
#define K_HALF 128
#define K_CLAMP 1048576.0f

typedef struct {
  int X, Y;
  float iz;
  int behind;
} vtx_t;

typedef struct {
  int ax, ay, bx, by, cx, cy;
  float iza, izb, izc;
  float inv_area;
  int flip;
  int ok;
} tri_t;

static int64_t edge_fn(int ax, int ay, int bx, int by, int px, int py) {
  return (int64_t)(bx - ax) * (int64_t)(py - ay) - (int64_t)(by - ay) * (int64_t)(px - ax);
}

static int floor_div(int a, int b) {
  int q = a / b;
  if ((a % b != 0) && (a < 0)) --q;
  return q;
}

static float normal_texture(float s) {
  const float u = fmaf(s, 32.0f, -0.5f);
  const float fl = floorf(u);
  const float f = u - fl;
  const int k0 = ((int)fl) & 31;
  const int k1 = (k0 + 1) & 31;
  const float t0 = (float)((k0 * 255) >> 5) / 255.0f; /* texel uint8(k*255/32) read back as /255 */
  const float t1 = (float)((k1 * 255) >> 5) / 255.0f;
  return fmaf(f, t1 - t0, t0);
}

static float quant8(float v, int on) {
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  if (!on) return v;
  return (float)lrintf(v * 255.0f) / 255.0f; /* uint8 level k, read back as k / 255 */
}

/* optional texture of one mesh: uv [nv,2], tex [th,tw,3] uint8 (row 0 = top of the image), modulate = multiply with the
 * interpolated vertex colours */
typedef struct {
  const float* uv;
  const uint8_t* tex;
  int th, tw;
  int modulate;
} tex_t;

static int wrap_idx(int i, int n) {
  int r = i % n;
  return r < 0 ? r + n : r;
}

static void texture_sample(const tex_t* t, float u, float v, float* out3) {
  if (!(u == u)) u = 0.f;
  if (!(v == v)) v = 0.f;
  u = u - floorf(u);
  v = v - floorf(v);
  const float x = fmaf(u, (float)t->tw, -0.5f);
  const float y = fmaf(1.0f - v, (float)t->th, -0.5f);
  const float x0 = floorf(x), y0 = floorf(y);
  const float fx = x - x0, fy = y - y0;
  const int i0 = wrap_idx((int)x0, t->tw), i1 = wrap_idx((int)x0 + 1, t->tw);
  const int r0 = wrap_idx((int)y0, t->th), r1 = wrap_idx((int)y0 + 1, t->th);
  for (int k = 0; k < 3; ++k) {
    const float c00 = (float)t->tex[((size_t)r0 * t->tw + i0) * 3 + k] / 255.0f;
    const float c10 = (float)t->tex[((size_t)r0 * t->tw + i1) * 3 + k] / 255.0f;
    const float c01 = (float)t->tex[((size_t)r1 * t->tw + i0) * 3 + k] / 255.0f;
    const float c11 = (float)t->tex[((size_t)r1 * t->tw + i1) * 3 + k] / 255.0f;
    const float top = fmaf(fx, c10 - c00, c00);
    const float bot = fmaf(fx, c11 - c01, c01);
    out3[k] = fmaf(fy, bot - top, top);
  }
}

static tri_t load_tri(const vtx_t* vtx, const int32_t* faces, int tri) {
  tri_t t;
  const vtx_t a = vtx[faces[3 * tri]], b = vtx[faces[3 * tri + 1]], c = vtx[faces[3 * tri + 2]];
  t.ax = a.X; t.ay = a.Y; t.bx = b.X; t.by = b.Y; t.cx = c.X; t.cy = c.Y;
  t.iza = a.iz; t.izb = b.iz; t.izc = c.iz;
  int64_t area2 = edge_fn(t.ax, t.ay, t.bx, t.by, t.cx, t.cy);
  t.flip = area2 < 0;
  if (t.flip) area2 = -area2;
  t.ok = (area2 != 0) && !(a.behind | b.behind | c.behind);
  t.inv_area = t.ok ? 1.0f / (float)area2 : 0.f;
  return t;
}

/* returns 1 and the barycentrics / interpolated 1/z if pixel centre (px, py) is covered and inside the depth range */
static int tri_sample(const tri_t* t, int px, int py, float* l0, float* l1, float* l2, float* iz) {
  int64_t w0 = edge_fn(t->bx, t->by, t->cx, t->cy, px, py);
  int64_t w1 = edge_fn(t->cx, t->cy, t->ax, t->ay, px, py);
  int64_t w2 = edge_fn(t->ax, t->ay, t->bx, t->by, px, py);
  if (t->flip) { w0 = -w0; w1 = -w1; w2 = -w2; }
  if ((w0 | w1 | w2) < 0) return 0;
  *l0 = (float)w0 * t->inv_area;
  *l1 = (float)w1 * t->inv_area;
  *l2 = (float)w2 * t->inv_area;
  *iz = fmaf(*l0, t->iza, fmaf(*l1, t->izb, *l2 * t->izc));
  return (*iz >= K_IZ_MIN) && (*iz <= K_IZ_MAX);
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/*
 * Render one view.  verts/normals/colors: [nv,3]; faces: [nf,3]; TCO row-major 4x4; K row-major 3x3.
 * flags: bit0 quantize8, bit1 GL eye axes.  Outputs (any may be NULL): rgb [3,h,w], nrm [3,h,w],
 * depth [h,w], tri_id [h,w] (-1 = background).  Returns 0, or -1 on allocation failure.
 */
static int render_view(const float* verts, const float* normals, const float* colors, int nv,
                       const int32_t* faces, int nf, const float* TCO, const float* K, int h, int w,
                       unsigned flags, const tex_t* texture, float* rgb, float* nrm, float* depth, int32_t* tri_id) {
  const int npix = h * w;
  const int q8 = (flags & 1u) != 0, gl_axes = (flags & 2u) != 0, lights = (flags & 4u) != 0;
  float light_dist = 0.f;
  if (lights) {
    float radius = 0.f;
    for (int i = 0; i < nv; ++i) {
      const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
      const float r = sqrtf(fmaf(x, x, fmaf(y, y, z * z)));
      if (r > radius) radius = r;
    }
    light_dist = radius * 10.0f;
  }
  int valid = 1;
  for (int i = 0; i < 16; ++i) valid = valid && isfinite(TCO[i]);
  for (int i = 0; i < 9; ++i) valid = valid && isfinite(K[i]);
  if (rgb) memset(rgb, 0, sizeof(float) * 3 * npix);
  if (nrm) memset(nrm, 0, sizeof(float) * 3 * npix);
  if (depth) memset(depth, 0, sizeof(float) * npix);
  if (tri_id) for (int i = 0; i < npix; ++i) tri_id[i] = -1;
  if (!valid) return 0;

  uint64_t* vis = (uint64_t*)malloc(sizeof(uint64_t) * npix);
  vtx_t* vtx = (vtx_t*)malloc(sizeof(vtx_t) * (nv > 0 ? nv : 1));
  if (!vis || !vtx) { free(vis); free(vtx); return -1; }
  for (int i = 0; i < npix; ++i) vis[i] = ~(uint64_t)0;
  const float* R = TCO;
  const float fx = K[0], cx = K[2], fy = K[4], cy = K[5];
  for (int i = 0; i < nv; ++i) {
    const float px = verts[3 * i], py = verts[3 * i + 1], pz = verts[3 * i + 2];
    const float xc = fmaf(R[0], px, fmaf(R[1], py, fmaf(R[2], pz, R[3])));
    const float yc = fmaf(R[4], px, fmaf(R[5], py, fmaf(R[6], pz, R[7])));
    const float zc = fmaf(R[8], px, fmaf(R[9], py, fmaf(R[10], pz, R[11])));
    vtx_t o;
    o.behind = !(zc >= K_PROJ_MIN);
    const float zs = o.behind ? 1.0f : zc;
    const float iz = 1.0f / zs;
    float u = fmaf(fx, xc * iz, cx);
    float v = fmaf(fy, yc * iz, cy);
    u = fminf(fmaxf(u, -K_CLAMP), K_CLAMP);
    v = fminf(fmaxf(v, -K_CLAMP), K_CLAMP);
    if (!(u == u)) { u = 0.f; o.behind = 1; }
    if (!(v == v)) { v = 0.f; o.behind = 1; }
    o.X = (int)lrintf(u * (float)K_SUB);
    o.Y = (int)lrintf(v * (float)K_SUB);
    o.iz = iz;
    vtx[i] = o;
  }
  for (int tri = 0; tri < nf; ++tri) {
    const tri_t t = load_tri(vtx, faces, tri);
    if (!t.ok) continue;
    const int minx = imin(t.ax, imin(t.bx, t.cx)), maxx = imax(t.ax, imax(t.bx, t.cx));
    const int miny = imin(t.ay, imin(t.by, t.cy)), maxy = imax(t.ay, imax(t.by, t.cy));
    const int j0 = imax(0, -floor_div(-(minx - K_HALF), K_SUB));
    const int j1 = imin(w - 1, floor_div(maxx - K_HALF, K_SUB));
    const int i0 = imax(0, -floor_div(-(miny - K_HALF), K_SUB));
    const int i1 = imin(h - 1, floor_div(maxy - K_HALF, K_SUB));
    for (int i = i0; i <= i1; ++i) {
      for (int j = j0; j <= j1; ++j) {
        float l0, l1, l2, iz;
        if (!tri_sample(&t, j * K_SUB + K_HALF, i * K_SUB + K_HALF, &l0, &l1, &l2, &iz)) continue;
        uint32_t zb;
        memcpy(&zb, &iz, 4);
        const uint64_t key = ((uint64_t)(~zb) << 32) | (uint32_t)tri;
        if (key < vis[i * w + j]) vis[i * w + j] = key;
      }
    }
  }
  const float dep_a = -0.10101010f;
  const float dep_b = 1.01010101f;
  for (int pix = 0; pix < npix; ++pix) {
    const uint64_t key = vis[pix];
    if (key == ~(uint64_t)0) continue;
    const int i = pix / w, j = pix - i * w;
    const int tri = (int)(key & 0xffffffffu);
    const tri_t t = load_tri(vtx, faces, tri);
    float l0, l1, l2, iz;
    tri_sample(&t, j * K_SUB + K_HALF, i * K_SUB + K_HALF, &l0, &l1, &l2, &iz);
    const float z = 1.0f / iz;
    const float b0 = (l0 * t.iza) * z;
    const float b1 = (l1 * t.izb) * z;
    const float b2 = (l2 * t.izc) * z;
    const int ia = faces[3 * tri], ib = faces[3 * tri + 1], ic = faces[3 * tri + 2];
    float col[3], nn[3];
    for (int k = 0; k < 3; ++k) {
      col[k] = fmaf(b0, colors[3 * ia + k], fmaf(b1, colors[3 * ib + k], b2 * colors[3 * ic + k]));
      nn[k] = fmaf(b0, normals[3 * ia + k], fmaf(b1, normals[3 * ib + k], b2 * normals[3 * ic + k]));
    }
    if (texture && texture->tex) {
      const float* uv = texture->uv;
      const float tu = fmaf(b0, uv[2 * ia], fmaf(b1, uv[2 * ib], b2 * uv[2 * ic]));
      const float tv = fmaf(b0, uv[2 * ia + 1], fmaf(b1, uv[2 * ib + 1], b2 * uv[2 * ic + 1]));
      float tc[3];
      texture_sample(texture, tu, tv, tc);
      for (int k = 0; k < 3; ++k) col[k] = texture->modulate ? tc[k] * col[k] : tc[k];
    }
    if (lights) {
      float p[3];
      for (int k = 0; k < 3; ++k)
        p[k] = fmaf(b0, verts[3 * ia + k], fmaf(b1, verts[3 * ib + k], b2 * verts[3 * ic + k]));
      float u0 = nn[0], u1 = nn[1], u2 = nn[2];
      const float ul = sqrtf(fmaf(u0, u0, fmaf(u1, u1, u2 * u2)));
      if (ul > 0.f) {
        const float inv = 1.0f / ul;
        u0 = u0 * inv; u1 = u1 * inv; u2 = u2 * inv;
      }
      float shade = 0.1f;
      for (int li = 0; li < 6; ++li) {
        const float sgn = (li & 1) ? -light_dist : light_dist;
        const float d0 = (li < 2 ? sgn : 0.f) - p[0];
        const float d1 = ((li >> 1) == 1 ? sgn : 0.f) - p[1];
        const float d2 = (li >= 4 ? sgn : 0.f) - p[2];
        const float dist = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
        if (dist > 0.f) {
          const float ndl = fmaf(u0, d0, fmaf(u1, d1, u2 * d2)) / dist;
          shade = fmaf(0.4f, fmaxf(ndl, 0.0f), shade);
        }
      }
      for (int k = 0; k < 3; ++k) col[k] = col[k] * shade;
    }
    if (rgb) {
      rgb[pix] = quant8(col[0], q8);
      rgb[npix + pix] = quant8(col[1], q8);
      rgb[2 * npix + pix] = quant8(col[2], q8);
    }
    if (nrm) {
      float ex = fmaf(R[0], nn[0], fmaf(R[1], nn[1], R[2] * nn[2]));
      float ey = fmaf(R[4], nn[0], fmaf(R[5], nn[1], R[6] * nn[2]));
      float ez = fmaf(R[8], nn[0], fmaf(R[9], nn[1], R[10] * nn[2]));
      const float len = sqrtf(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
      if (len > 0.f) {
        const float inv = 1.0f / len;
        ex = ex * inv; ey = ey * inv; ez = ez * inv;
      }
      const float px_ = ex;
      const float py_ = gl_axes ? -ey : ez;
      const float pz_ = gl_axes ? -ez : -ey;
      nrm[pix] = quant8(normal_texture(px_), q8);
      nrm[npix + pix] = quant8(normal_texture(py_), q8);
      nrm[2 * npix + pix] = quant8(normal_texture(pz_), q8);
    }
    if (depth) {
      const float d = fmaf(dep_a, iz, dep_b);
      depth[pix] = (d > 0.999f) ? 0.f : z;
    }
    if (tri_id) tri_id[pix] = tri;
  }
  free(vis);
  free(vtx);
  return 0;
}

int raster_ref_render(const float* verts, const float* normals, const float* colors, int nv,
                      const int32_t* faces, int nf, const float* TCO, const float* K, int h, int w,
                      unsigned flags, float* rgb, float* nrm, float* depth, int32_t* tri_id) {
  return render_view(verts, normals, colors, nv, faces, nf, TCO, K, h, w, flags, NULL, rgb, nrm, depth, tri_id);
}

/* the same with a texture: uv [nv,2], tex [th,tw,3] uint8 */
int raster_ref_render_tex(const float* verts, const float* normals, const float* colors, const float* uv, int nv,
                          const int32_t* faces, int nf, const uint8_t* tex, int th, int tw, int modulate,
                          const float* TCO, const float* K, int h, int w, unsigned flags, float* rgb, float* nrm,
                          float* depth, int32_t* tri_id) {
  tex_t t;
  t.uv = uv; t.tex = tex; t.th = th; t.tw = tw; t.modulate = modulate;
  return render_view(verts, normals, colors, nv, faces, nf, TCO, K, h, w, flags, (tex && uv) ? &t : NULL, rgb, nrm,
                     depth, tri_id);
}

/* Batched convenience: n views of (possibly different) meshes given by offsets, as the device API.
 * Views are spread over n_threads POSIX threads (<= 0: one thread). */
#include <pthread.h>

typedef struct {
  const float *verts, *normals, *colors;
  const int64_t *vert_offsets, *face_offsets;
  const int32_t *faces, *label_idx;
  const float *TCO, *K;
  int n_views, h, w;
  unsigned flags;
  float *rgb, *nrm, *depth;
  /* optional textures: uv [sum_nv,2]; tex = all textures back to back; per mesh byte offset, (th, tw), modulate flag */
  const float* uv;
  const uint8_t* tex;
  const int64_t* tex_offsets;
  const int32_t* tex_dims;
  const int32_t* tex_modulate;
  int next;  /* work counter, protected by lock */
  int status;
  pthread_mutex_t lock;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  const size_t npix = (size_t)b->h * b->w;
  for (;;) {
    pthread_mutex_lock(&b->lock);
    const int v = b->next++;
    pthread_mutex_unlock(&b->lock);
    if (v >= b->n_views) break;
    const int lab = b->label_idx[v];
    const int64_t vo = b->vert_offsets[lab], fo = b->face_offsets[lab];
    tex_t t;
    const tex_t* tp = NULL;
    if (b->tex && b->uv && b->tex_dims[2 * lab] > 0 && b->tex_dims[2 * lab + 1] > 0) {
      t.uv = b->uv + 2 * vo;
      t.tex = b->tex + b->tex_offsets[lab];
      t.th = b->tex_dims[2 * lab];
      t.tw = b->tex_dims[2 * lab + 1];
      t.modulate = b->tex_modulate ? b->tex_modulate[lab] : 0;
      tp = &t;
    }
    const int rc = render_view(b->verts + 3 * vo, b->normals + 3 * vo, b->colors + 3 * vo,
                               (int)(b->vert_offsets[lab + 1] - vo), b->faces + 3 * fo,
                               (int)(b->face_offsets[lab + 1] - fo), b->TCO + 16 * v, b->K + 9 * v,
                               b->h, b->w, b->flags, tp, b->rgb ? b->rgb + 3 * npix * v : NULL,
                               b->nrm ? b->nrm + 3 * npix * v : NULL,
                               b->depth ? b->depth + npix * v : NULL, NULL);
    if (rc != 0) {
      pthread_mutex_lock(&b->lock);
      b->status = rc;
      pthread_mutex_unlock(&b->lock);
    }
  }
  return NULL;
}

int raster_ref_render_batch_tex(int n_meshes, const float* verts, const float* normals, const float* colors,
                                const int64_t* vert_offsets, const int32_t* faces, const int64_t* face_offsets,
                                const float* uv, const uint8_t* tex, const int64_t* tex_offsets,
                                const int32_t* tex_dims, const int32_t* tex_modulate, const int32_t* label_idx,
                                const float* TCO, const float* K, int n_views, int h, int w, unsigned flags,
                                float* rgb, float* nrm, float* depth, int n_threads);

int raster_ref_render_batch(int n_meshes, const float* verts, const float* normals, const float* colors,
                            const int64_t* vert_offsets, const int32_t* faces, const int64_t* face_offsets,
                            const int32_t* label_idx, const float* TCO, const float* K, int n_views, int h,
                            int w, unsigned flags, float* rgb, float* nrm, float* depth, int n_threads) {
  return raster_ref_render_batch_tex(n_meshes, verts, normals, colors, vert_offsets, faces, face_offsets, NULL, NULL,
                                     NULL, NULL, NULL, label_idx, TCO, K, n_views, h, w, flags, rgb, nrm, depth,
                                     n_threads);
}

int raster_ref_render_batch_tex(int n_meshes, const float* verts, const float* normals, const float* colors,
                                const int64_t* vert_offsets, const int32_t* faces, const int64_t* face_offsets,
                                const float* uv, const uint8_t* tex, const int64_t* tex_offsets,
                                const int32_t* tex_dims, const int32_t* tex_modulate, const int32_t* label_idx,
                                const float* TCO, const float* K, int n_views, int h, int w, unsigned flags,
                                float* rgb, float* nrm, float* depth, int n_threads) {
  for (int v = 0; v < n_views; ++v)
    if (label_idx[v] < 0 || label_idx[v] >= n_meshes) return -2;
  batch_t b;
  b.uv = uv; b.tex = tex; b.tex_offsets = tex_offsets; b.tex_dims = tex_dims; b.tex_modulate = tex_modulate;
  b.verts = verts; b.normals = normals; b.colors = colors;
  b.vert_offsets = vert_offsets; b.face_offsets = face_offsets;
  b.faces = faces; b.label_idx = label_idx; b.TCO = TCO; b.K = K;
  b.n_views = n_views; b.h = h; b.w = w; b.flags = flags;
  b.rgb = rgb; b.nrm = nrm; b.depth = depth;
  b.next = 0; b.status = 0;
  pthread_mutex_init(&b.lock, NULL);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if (n_threads > n_views) n_threads = n_views > 0 ? n_views : 1;
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < n_threads - 1; ++i)
    if (pthread_create(&th[started], NULL, batch_worker, &b) == 0) ++started;
  batch_worker(&b);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  pthread_mutex_destroy(&b.lock);
  return b.status;
}
