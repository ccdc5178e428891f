// raster.cuh -- the per-tile item-buffer rasteriser shared by both K4 variants (visibility.cu: BVH frustum packets per tile;
// visibility_bins.cu: depth-sorted leaf bins per 128x128 supertile).  One warp owns one 32x32-pixel tile: nearest hit per pixel
// centre in shared memory, up to 32 triangles set up in parallel (one per lane, fp64) and tested one after the other on the 8x4
// blocks that can still matter.  Replaces viewer.frame() + glReadPixels + getAllColorsInFrame
// (ImageViewerCaptureTool.cpp:59-70, CameraEstimator.cpp:99-129).
//
// Numerics that matter for parity:
//  * every edge normal is evaluated from the lexicographically lower endpoint of its edge (vertices are stored sorted), so two
//    triangles sharing an edge get bit-identical coefficients up to sign and no pixel ray slips between them;
//  * depth resolution: fragments within a relative 2^-16 of the nearest are ties and the lowest id wins (GL_LESS on a
//    fixed-point depth buffer, triangles drawn in id order).
#pragma once
#include "kernels.h"

namespace ipp {
namespace raster {

constexpr int kTile = 32;  // tile side in pixels
constexpr int kTilePixels = kTile * kTile;
constexpr int kNoId = 0x7fffffff;
constexpr float kDepthTie = 1.0f + 1.0f / 65536.0f;
constexpr int kTriWords = 16;  // floats per staged triangle
constexpr int kPending = 32;   // triangles staged together (one per lane)
#ifndef IPP_BLOCKS_PER_STEP
#define IPP_BLOCKS_PER_STEP 2
#endif
constexpr int kBlocksPerStep = IPP_BLOCKS_PER_STEP;  // 8x4 blocks one step of the raster loop tests

// a*b - c*d and a.b with a fixed rounding sequence.  Explicit intrinsics: the compiler may neither contract nor reassociate them, so
// the same inputs give the same bits everywhere and negated inputs give exactly negated results -- the property the shared-edge
// argument rests on (-(x*y - z*w) silently rewritten as fma(z, w, -(x*y)) rounds the other product first and breaks it).
__device__ __forceinline__ double ddiff(double a, double b, double c, double d) { return __fma_rn(a, b, -__dmul_rn(c, d)); }
__device__ __forceinline__ double ddot(double ax, double ay, double az, double bx, double by, double bz) {
    return __fma_rn(az, bz, __fma_rn(ay, by, __dmul_rn(ax, bx)));
}

// MUFU.RCP without the range scaling __fdividef wraps around it (4 instructions per division): a denormal determinant flushes to
// zero, the depth becomes inf / NaN and fails the range test like an exactly grazing ray does.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per-tile constants of one lane: pixel-centre image-plane coordinates X(px) = ((px + 0.5) / W * 2 - 1) tanX = px kx + cx.
struct TileGeom {
    float kx8, ky4;      // image-plane step between 8x4 blocks
    float Xl, Yl;        // this lane's pixel in block (0, 0); block (bx, by) adds (bx kx8, by ky4)
    float tXmin, tXmax, tYmin, tYmax;  // extreme pixel centres of the tile
    float bXmin, bXmax, bYmin, bYmax;  // extreme pixel centres of block b = lane
    int sox = 0, soy = 0;              // the tile's first pixel inside its 128x128 supertile (binned variant)
};

// The same expression is used for a pixel whatever triangle is tested, so shared edges stay watertight; the tile and block
// extremes use the expressions their corner pixels use.
__device__ __forceinline__ void tile_geometry(TileGeom& g, int tpx0, int tpy0, int lane, float kx, float cx, float ky, float cy) {
    g.kx8 = 8.f * kx;
    g.ky4 = 4.f * ky;
    g.Xl = __fmaf_rn((float)(tpx0 + (lane & 7)), kx, cx);
    g.Yl = __fmaf_rn((float)(tpy0 + (lane >> 3)), ky, cy);
    g.tXmin = __fmaf_rn((float)tpx0, kx, cx);
    g.tXmax = __fmaf_rn(3.f, g.kx8, __fmaf_rn((float)(tpx0 + 7), kx, cx));
    g.tYmin = __fmaf_rn((float)tpy0, ky, cy);
    g.tYmax = __fmaf_rn(7.f, g.ky4, __fmaf_rn((float)(tpy0 + 3), ky, cy));
    g.bXmin = __fmaf_rn((float)(lane & 3), g.kx8, __fmaf_rn((float)tpx0, kx, cx));
    g.bXmax = __fmaf_rn((float)(lane & 3), g.kx8, __fmaf_rn((float)(tpx0 + 7), kx, cx));
    g.bYmin = __fmaf_rn((float)(lane >> 2), g.ky4, __fmaf_rn((float)tpy0, ky, cy));
    g.bYmax = __fmaf_rn((float)(lane >> 2), g.ky4, __fmaf_rn((float)(tpy0 + 3), ky, cy));
}

// hit buffers of a tile: index = block * 32 + lane (block = 8x4 pixels).  Rays outside the image start with depth 0: no fragment
// can ever beat it.
__device__ __forceinline__ void tile_init(float* __restrict__ St, int* __restrict__ Sid, int lane, int vw, int vh, float zFar) {
#pragma unroll
    for (int k = 0; k < kTilePixels / 32; ++k) {
        const bool inside = ((k & 3) * 8 + (lane & 7)) < vw && ((k >> 2) * 4 + (lane >> 3)) < vh;
        St[k * 32 + lane] = inside ? zFar : 0.f;
        Sid[k * 32 + lane] = kNoId;
    }
}

// One lane sets up one triangle for the current camera pose: 16 floats in shared memory --
//   [0..2] [3..5] [6..8] edge functions e(X, Y) = o[0] + o[1] X + o[2] Y of the pixel ray D = f + X s + Y u, [9..11] the same for
//   det = D.Ng, [12] vol = (A - eye).Ng, [13] triangle id, [15] nearest vertex depth (< 0: produces no fragment in this tile).
// Everything is multiplied by sign(vol), so that a ray in front of the eye crosses the triangle iff all three edge functions are
// >= 0 and det > 0.  Setup runs in fp64 from the exact fp32 vertices, the fp32 eye and the fp64 camera basis; only the final
// coefficients are rounded to fp32 (in fp32 the normal of an edge that points at the camera, |A' x e| << |A'||e|, lost up to
// 0.03 pixel).
// kPoseOnly: the part that does not depend on the tile -- zcull is then the far bound of the pose, [14] receives the nearest vertex
// depth itself and the tile test is left to stage_record() below, which repeats it with the same expressions.
// It also reports, in *tilesOut, which of the 4 x 4 tiles of the 128 x 128 supertile `super` (bit = row * 4 + column) hold a pixel centre
// inside the screen bounding box of the projected vertices widened by kBoxPadPx pixels (all 16 when a vertex is not in front of the
// eye): a pixel centre that passes the three edge tests lies inside the projected triangle up to rounding << kBoxPadPx, so a tile
// outside the mask gets no fragment of the triangle.  *tilesOut is left alone when the triangle is rejected.
struct SuperGeom {
    float X0, Y0;    // image-plane coordinates of the supertile's first pixel centre
    float ipx, ipy;  // pixels per image-plane unit
    // extreme pixel centres of the supertile's four tile columns ([c] = min, [4 + c] = max) and rows, the very values
    // tile_geometry() gives a tile (tXmin, tXmax, tYmin, tYmax); nullptr: the screen-box mask is not refined
    const float* tX = nullptr;
    const float* tY = nullptr;
};
#ifndef IPP_EXACT_TILE_MASK
#define IPP_EXACT_TILE_MASK 1
#endif
constexpr float kBoxPadPx = 0.5f;
template <bool kPoseOnly>
__device__ __forceinline__ bool setup_triangle_t(const DeviceScene& sc, const Snapshot* sp, int tri, float* __restrict__ o, float ox,
                                                 float oy, float oz, float zNear, float zcull, const TileGeom& g,
                                                 const SuperGeom* super = nullptr, unsigned* tilesOut = nullptr) {
    const float4 a = __ldg(sc.ta + tri), b = __ldg(sc.tb + tri), cc = __ldg(sc.tc + tri);
    float valid = -1.f;
    unsigned boxWord = 0u;  // kPoseOnly: screen box of the triangle in pixels of the supertile
    if (b.w == 0.f) {  // not degenerate
        const double fX = sp->f[0], fY = sp->f[1], fZ = sp->f[2];
        const double ax = (double)a.x - (double)ox, ay = (double)a.y - (double)oy, az = (double)a.z - (double)oz;
        const double bx_ = (double)b.x - (double)ox, by_ = (double)b.y - (double)oy, bz_ = (double)b.z - (double)oz;
        const double cx_ = (double)cc.x - (double)ox, cy_ = (double)cc.y - (double)oy, cz_ = (double)cc.z - (double)oz;
        // camera-space depth of the vertices; the depth of a fragment is a convex combination of them
        const float za = (float)ddot(ax, ay, az, fX, fY, fZ), zb = (float)ddot(bx_, by_, bz_, fX, fY, fZ),
                    zc = (float)ddot(cx_, cy_, cz_, fX, fY, fZ);
        const float zlo = fminf(za, fminf(zb, zc)), zhi = fmaxf(za, fmaxf(zb, zc));
        if (zhi >= zNear * 0.9999f && zlo <= zcull * 1.000001f) {
            const double sX = sp->s[0], sY = sp->s[1], sZ = sp->s[2];
            const double uX = sp->u[0], uY = sp->u[1], uZ = sp->u[2];
            const double ebcx = (double)cc.x - (double)b.x, ebcy = (double)cc.y - (double)b.y, ebcz = (double)cc.z - (double)b.z;
            const double eacx = (double)cc.x - (double)a.x, eacy = (double)cc.y - (double)a.y, eacz = (double)cc.z - (double)a.z;
            const double eabx = (double)b.x - (double)a.x, eaby = (double)b.y - (double)a.y, eabz = (double)b.z - (double)a.z;
            // B' x (C-B);  C' x (A-C) = -(A' x (C-A));  A' x (B-A);  Ng = (B-A) x (C-A): every edge normal from the lower endpoint
            const double nax = ddiff(by_, ebcz, bz_, ebcy), nay = ddiff(bz_, ebcx, bx_, ebcz), naz = ddiff(bx_, ebcy, by_, ebcx);
            const double nbx = -ddiff(ay, eacz, az, eacy), nby = -ddiff(az, eacx, ax, eacz), nbz = -ddiff(ax, eacy, ay, eacx);
            const double ncx = ddiff(ay, eabz, az, eaby), ncy = ddiff(az, eabx, ax, eabz), ncz = ddiff(ax, eaby, ay, eabx);
            const double ngx = ddiff(eaby, eacz, eabz, eacy), ngy = ddiff(eabz, eacx, eabx, eacz), ngz = ddiff(eabx, eacy, eaby, eacx);
            const double vol = ddot(ax, ay, az, ngx, ngy, ngz);  // (A - eye).Ng
            const float sg = vol < 0.0 ? -1.f : 1.f;             // exact sign flip: shared edges stay bit-identical up to sign
            valid = fmaxf(zlo, 0.f) * 0.999999f;
            o[0] = sg * (float)ddot(nax, nay, naz, fX, fY, fZ); o[1] = sg * (float)ddot(nax, nay, naz, sX, sY, sZ); o[2] = sg * (float)ddot(nax, nay, naz, uX, uY, uZ);
            o[3] = sg * (float)ddot(nbx, nby, nbz, fX, fY, fZ); o[4] = sg * (float)ddot(nbx, nby, nbz, sX, sY, sZ); o[5] = sg * (float)ddot(nbx, nby, nbz, uX, uY, uZ);
            o[6] = sg * (float)ddot(ncx, ncy, ncz, fX, fY, fZ); o[7] = sg * (float)ddot(ncx, ncy, ncz, sX, sY, sZ); o[8] = sg * (float)ddot(ncx, ncy, ncz, uX, uY, uZ);
            o[9] = sg * (float)ddot(ngx, ngy, ngz, fX, fY, fZ); o[10] = sg * (float)ddot(ngx, ngy, ngz, sX, sY, sZ); o[11] = sg * (float)ddot(ngx, ngy, ngz, uX, uY, uZ);
            o[12] = sg * (float)vol;
            o[13] = a.w;  // triangle id (int bits)
            if constexpr (kPoseOnly) {
                o[14] = zlo;
                unsigned tiles = 0xffffu;
                boxWord = 127u << 7 | 127u << 21;  // a vertex is not in front of the eye: the whole supertile
                if (zlo > 1e-4f) {
                    const float ra = rcp_approx(za), rb = rcp_approx(zb), rc = rcp_approx(zc);
                    const float Xa = (float)ddot(ax, ay, az, sX, sY, sZ) * ra, Xb = (float)ddot(bx_, by_, bz_, sX, sY, sZ) * rb,
                                Xc = (float)ddot(cx_, cy_, cz_, sX, sY, sZ) * rc;
                    const float Ya = (float)ddot(ax, ay, az, uX, uY, uZ) * ra, Yb = (float)ddot(bx_, by_, bz_, uX, uY, uZ) * rb,
                                Yc = (float)ddot(cx_, cy_, cz_, uX, uY, uZ) * rc;
                    // in pixels from the supertile's first pixel centre: tile column c holds the centres 32 c .. 32 c + 31
                    const float fx0 = fmaxf((fminf(Xa, fminf(Xb, Xc)) - super->X0) * super->ipx - kBoxPadPx, -512.f);
                    const float fx1 = fminf((fmaxf(Xa, fmaxf(Xb, Xc)) - super->X0) * super->ipx + kBoxPadPx, 512.f);
                    const float fy0 = fmaxf((fminf(Ya, fminf(Yb, Yc)) - super->Y0) * super->ipy - kBoxPadPx, -512.f);
                    const float fy1 = fminf((fmaxf(Ya, fmaxf(Yb, Yc)) - super->Y0) * super->ipy + kBoxPadPx, 512.f);
                    const int c0 = max(0, (int)ceilf((fx0 - 31.f) * 0.03125f)), c1 = min(3, (int)floorf(fx1 * 0.03125f));
                    const int r0 = max(0, (int)ceilf((fy0 - 31.f) * 0.03125f)), r1 = min(3, (int)floorf(fy1 * 0.03125f));
                    tiles = 0u;
                    if (c0 <= c1 && r0 <= r1)
                        tiles = (((2u << c1) - (1u << c0)) * 0x1111u) & ((16u << (4 * r1)) - 1u) & ~((1u << (4 * r0)) - 1u);
                    // the pixel columns and rows of the supertile with a centre inside that box (7 bits each): stage_record() turns
                    // them into the 8x4 blocks of a tile the triangle can touch
                    const int px0 = min(127, max(0, (int)ceilf(fx0))), px1 = min(127, max(0, (int)floorf(fx1)));
                    const int py0 = min(127, max(0, (int)ceilf(fy0))), py1 = min(127, max(0, (int)floorf(fy1)));
                    boxWord = (unsigned)px0 | ((unsigned)px1 << 7) | ((unsigned)py0 << 14) | ((unsigned)py1 << 21);
                }
#if IPP_EXACT_TILE_MASK
                if (super->tX && tiles != 0u) {
                    // ... and of those, the tiles that pass the tile-level edge test stage_record() would apply (same expressions, same
                    // bits): a long thin triangle touches a fraction of the tiles of its screen box, and every tile left in the mask
                    // fetches the record, stages it and throws it away
                    unsigned exact = 0u;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (((tiles >> (4 * r)) & 0xfu) == 0u) continue;
                        const float Ymin = super->tY[r], Ymax = super->tY[4 + r];
                        const float ia = __fmaf_rn(o[2], o[2] >= 0.f ? Ymax : Ymin, o[0]);
                        const float ib = __fmaf_rn(o[5], o[5] >= 0.f ? Ymax : Ymin, o[3]);
                        const float ic = __fmaf_rn(o[8], o[8] >= 0.f ? Ymax : Ymin, o[6]);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float Xmin = super->tX[c], Xmax = super->tX[4 + c];
                            const float ma = __fmaf_rn(o[1], o[1] >= 0.f ? Xmax : Xmin, ia);
                            const float mb = __fmaf_rn(o[4], o[4] >= 0.f ? Xmax : Xmin, ib);
                            const float mc = __fmaf_rn(o[7], o[7] >= 0.f ? Xmax : Xmin, ic);
                            if (fminf(fminf(ma, mb), mc) >= 0.f) exact |= 1u << (4 * r + c);
                        }
                    }
                    tiles &= exact;
                }
#endif
                *tilesOut = tiles;
            } else {
            // can any pixel centre of the tile be inside?  Each edge function is monotone in X and in Y (also as computed), so its
            // value at the extreme pixel centre of the tile bounds it over the tile.
            const float ma = __fmaf_rn(o[1], o[1] >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(o[2], o[2] >= 0.f ? g.tYmax : g.tYmin, o[0]));
            const float mb = __fmaf_rn(o[4], o[4] >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(o[5], o[5] >= 0.f ? g.tYmax : g.tYmin, o[3]));
            const float mc = __fmaf_rn(o[7], o[7] >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(o[8], o[8] >= 0.f ? g.tYmax : g.tYmin, o[6]));
            if (fminf(fminf(ma, mb), mc) < 0.f) valid = -1.f;
            }
        }
    }
    // a record keeps the screen box where the staged triangle keeps `valid` (stage_record() derives that from [14] again); its sign
    // bit says whether the triangle survived
    if constexpr (kPoseOnly) o[15] = valid >= 0.f ? __uint_as_float(boxWord) : -1.f;
    else o[15] = valid;
    return valid >= 0.f;
}
__device__ __forceinline__ bool setup_triangle(const DeviceScene& sc, const Snapshot* sp, int tri, float* __restrict__ o, float ox,
                                               float oy, float oz, float zNear, float zcull, const TileGeom& g) {
    return setup_triangle_t<false>(sc, sp, tri, o, ox, oy, oz, zNear, zcull, g);
}

// The tile's half of the set-up for a triangle whose pose half (setup_triangle_t<true>) sits in a 64-byte record r: the depth test
// against the tile's farthest hit and the tile-level edge test, then the staged copy in shared memory.  Same expressions, same
// bits as setup_triangle().  r was written by this warp earlier in the same kernel: plain loads.
__device__ __forceinline__ bool stage_record(const float4* r, float* __restrict__ o, float zcull, const TileGeom& g) {
    const float4 q0 = r[0], q1 = r[1], q2 = r[2], q3 = r[3];
    float valid = __float_as_int(q3.w) >= 0 ? fmaxf(q3.z, 0.f) * 0.999999f : -1.f;  // as setup_triangle_t computes it
    unsigned blocks = 0u;
    if (valid >= 0.f && q3.z <= zcull * 1.000001f) {
        const float ma = __fmaf_rn(q0.y, q0.y >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(q0.z, q0.z >= 0.f ? g.tYmax : g.tYmin, q0.x));
        const float mb = __fmaf_rn(q1.x, q1.x >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(q1.y, q1.y >= 0.f ? g.tYmax : g.tYmin, q0.w));
        const float mc = __fmaf_rn(q1.w, q1.w >= 0.f ? g.tXmax : g.tXmin, __fmaf_rn(q2.x, q2.x >= 0.f ? g.tYmax : g.tYmin, q1.z));
        if (fminf(fminf(ma, mb), mc) < 0.f) valid = -1.f;
        // the 8x4 blocks of this tile (bit = row * 4 + column) that hold a pixel of the triangle's screen box: the three edge
        // functions alone also accept blocks beyond a small triangle's corners (and, for a sliver whose coefficients are all
        // rounding noise, pixels that are nowhere near it)
        const unsigned w = __float_as_uint(q3.w);
        const int x0 = (int)(w & 127u) - g.sox, x1 = (int)((w >> 7) & 127u) - g.sox;
        const int y0 = (int)((w >> 14) & 127u) - g.soy, y1 = (int)((w >> 21) & 127u) - g.soy;
        if (x1 >= 0 && y1 >= 0 && x0 < kTile && y0 < kTile) {
            const int bx0 = max(x0, 0) >> 3, bx1 = min(x1, kTile - 1) >> 3, by0 = max(y0, 0) >> 2, by1 = min(y1, kTile - 1) >> 2;
            const unsigned cols = (2u << bx1) - (1u << bx0);  // 4 bits
            const unsigned rows = (by1 >= 7 ? 0xffffffffu : ((16u << (4 * by1)) - 1u)) & ~((1u << (4 * by0)) - 1u);
            blocks = (cols * 0x11111111u) & rows;
        }
        if (blocks == 0u) valid = -1.f;
    } else {
        valid = -1.f;
    }
    float4* o4 = reinterpret_cast<float4*>(o);
    o4[0] = q0; o4[1] = q1; o4[2] = q2; o4[3] = make_float4(q3.x, q3.y, __uint_as_float(blocks), valid);
    return valid >= 0.f;
}

// Tests the staged triangles named by `live` (bit k = S.tri[k]) against the tile, nearest first come first served in staging
// order.  Lane b first answers for 8x4 block b (edge functions at the block's extreme pixel centres, the triangle's nearest vertex
// depth against the block's farthest hit bz): one ballot gives the blocks that can still matter; the warp then runs the
// ray/triangle test on those, every lane owning the same 32 pixels of the tile throughout.  Edge functions are monotone in X and Y
// as computed (explicit fmaf), so the block rejection is exact.  Returns true when some hit changed.
template <bool kBlockMask = false>
__device__ __forceinline__ bool raster_staged(unsigned live, const float (*__restrict__ tri)[kTriWords], float* __restrict__ St,
                                              int* __restrict__ Sid, const TileGeom& g, int lane, float zNear, float zFar, float zcull,
                                              float bz, unsigned* diagBlocks = nullptr) {
    bool dirty = false;
    while (live) {
        const int k = __ffs(live) - 1;
        live &= live - 1;
        const float* tr = tri[k];
        const float zloTri = tr[15];
        if (zloTri > zcull) continue;
        const float eaf = tr[0], eas = tr[1], eau = tr[2], ebf = tr[3], ebs = tr[4], ebu = tr[5], ecf = tr[6], ecs = tr[7], ecu = tr[8],
                    gf = tr[9], gs = tr[10], gu = tr[11], vol = tr[12];
        const int id = __float_as_int(tr[13]);
        const float ca = __fmaf_rn(eas, eas >= 0.f ? g.bXmax : g.bXmin, __fmaf_rn(eau, eau >= 0.f ? g.bYmax : g.bYmin, eaf));
        const float cb = __fmaf_rn(ebs, ebs >= 0.f ? g.bXmax : g.bXmin, __fmaf_rn(ebu, ebu >= 0.f ? g.bYmax : g.bYmin, ebf));
        const float cc = __fmaf_rn(ecs, ecs >= 0.f ? g.bXmax : g.bXmin, __fmaf_rn(ecu, ecu >= 0.f ? g.bYmax : g.bYmin, ecf));
        // nearest depth the triangle's plane can have inside block b: depth = vol / det and det is linear, so its largest value over
        // the block sits at an extreme pixel centre; pushed up by a bound on its rounding error (|X|, |Y| < 1) and compared with a
        // relative margin far above that of the per-pixel depth.  det <= 0 everywhere: no ray of the block meets the front of the plane.
        const float dmax = __fmaf_rn(gs, gs >= 0.f ? g.bXmax : g.bXmin, __fmaf_rn(gu, gu >= 0.f ? g.bYmax : g.bYmin, gf)) +
                           2e-6f * (fabsf(gs) + fabsf(gu) + fabsf(gf));
        const float zloBlk = fmaxf(zloTri, vol * rcp_approx(dmax) * 0.9999f);
        unsigned todo = __ballot_sync(0xffffffffu, fminf(fminf(ca, cb), cc) >= 0.f && dmax > 0.f && zloBlk <= bz);
        if constexpr (kBlockMask) todo &= __float_as_uint(tr[14]);  // blocks inside the triangle's screen box (stage_record)
        if (diagBlocks) *diagBlocks += __popc(todo);
        // kBlocksPerStep blocks per step: their tests are independent (different pixels), which hides the latency of one behind
        // the others
        while (todo) {
            int blk[kBlocksPerStep];
            bool use[kBlocksPerStep];
#pragma unroll
            for (int u = 0; u < kBlocksPerStep; ++u) {
                use[u] = todo != 0u;
                blk[u] = use[u] ? __ffs(todo) - 1 : 0;
                todo &= todo - 1;  // no-op on 0
            }
            float tt[kBlocksPerStep], ht[kBlocksPerStep];
            bool in[kBlocksPerStep];
            int idx[kBlocksPerStep], hid[kBlocksPerStep];
#pragma unroll
            for (int u = 0; u < kBlocksPerStep; ++u) {
                const float X = __fmaf_rn((float)(blk[u] & 3), g.kx8, g.Xl), Y = __fmaf_rn((float)(blk[u] >> 2), g.ky4, g.Yl);
                const float eu = __fmaf_rn(eas, X, __fmaf_rn(eau, Y, eaf)), ev = __fmaf_rn(ebs, X, __fmaf_rn(ebu, Y, ebf)),
                            ew = __fmaf_rn(ecs, X, __fmaf_rn(ecu, Y, ecf));
                in[u] = use[u] && fminf(fminf(eu, ev), ew) >= 0.f;  // the ray crosses the triangle
                tt[u] = vol * rcp_approx(gs * X + (gu * Y + gf));    // det == 0: inf / NaN fails the range test
                idx[u] = blk[u] * 32 + lane;
                ht[u] = St[idx[u]];
                hid[u] = Sid[idx[u]];  // read with the depth, not after the depth test: one shared-memory round trip per step
            }
#pragma unroll
            for (int u = 0; u < kBlocksPerStep; ++u) {
                // depth-tie band: fragments within a relative 2^-16 of the nearest are ties and the lowest id wins
                if (in[u] && tt[u] >= zNear && tt[u] <= zFar && tt[u] <= ht[u] * kDepthTie) {
                    Sid[idx[u]] = (tt[u] * kDepthTie < ht[u]) ? id : min(hid[u], id);  // clearly nearer: replace; inside the band: lowest id
                    St[idx[u]] = fminf(ht[u], tt[u]);
                    dirty = true;
                }
            }
        }
    }
    return dirty;
}

// Hierarchical depth culling after a change: lane b takes the farthest hit of block b (conflict-free diagonal walk); rays without
// a hit still hold zFar, rays outside the image hold 0.  bz = bound of block b = lane, zcull = bound of the tile.
__device__ __forceinline__ void refresh_depth_bounds(const float* __restrict__ St, int lane, float& bz, float& zcull) {
    float zm = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) zm = fmaxf(zm, St[lane * 32 + ((j + lane) & 31)]);
    zm *= kDepthTie;
    bz = zm;
    for (int o2 = 16; o2; o2 >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o2));
    zcull = zm;
}

// The winners: one byte store per run of equal ids inside an 8x4 block (flagRow[id] = 1); optional id image for the test hook.
__device__ __forceinline__ void tile_output_row(const int* __restrict__ Sid, int lane, uint8_t* __restrict__ flagRow, int32_t* __restrict__ idImage,
                                                int tpx0, int tpy0, int W, int H) {
#pragma unroll 4
    for (int k = 0; k < kTilePixels / 32; ++k) {
        const int id = Sid[k * 32 + lane];
        const int left = __shfl_up_sync(0xffffffffu, id, 1), up = __shfl_up_sync(0xffffffffu, id, 8);
        const bool dupe = ((lane & 7) != 0 && left == id) || (lane >= 8 && up == id);
        if (id != kNoId && !dupe) flagRow[id] = 1;
        if (idImage) {
            const int px = tpx0 + (k & 3) * 8 + (lane & 7), py = tpy0 + (k >> 2) * 4 + (lane >> 3);
            if (px < W && py < H) idImage[(size_t)py * W + px] = id == kNoId ? -1 : id;
        }
    }
}
// ... into the flag row of every edge that shares this camera pose (slots[0 .. nSlots), see pose_dedupe.cu: one row unless several
// edges meet at the pose)
__device__ __forceinline__ void tile_output(const int* __restrict__ Sid, int lane, uint8_t* __restrict__ flags, size_t Tpad,
                                            const int* __restrict__ slots, int nSlots, int32_t* __restrict__ idImage, int tpx0, int tpy0, int W,
                                            int H) {
    tile_output_row(Sid, lane, flags + (size_t)__ldg(slots) * Tpad, idImage, tpx0, tpy0, W, H);
#pragma unroll 1
    for (int q = 1; q < nSlots; ++q) tile_output_row(Sid, lane, flags + (size_t)__ldg(slots + q) * Tpad, nullptr, tpx0, tpy0, W, H);
}

// Work-list lookup of the persistent kernels.
// last index s in [0, nSnaps) with tileOffset[s] <= task (tileOffset[0] == 0): 32-ary search, one probe per lane and round
__device__ __forceinline__ int find_snapshot(const int* __restrict__ tileOffset, int nSnaps, long long task, int lane) {
    int lo = 0, hi = nSnaps - 1;
    while (hi > lo) {
        int span = hi - lo;
        int step = (span + 31) >> 5;
        int pos = min(hi, lo + (lane + 1) * step);
        bool le = (long long)__ldg(tileOffset + pos) <= task;
        int k = __popc(__ballot_sync(0xffffffffu, le));
        int nlo = k ? min(hi, lo + k * step) : lo;
        int nhi = (k < 32) ? min(hi, lo + (k + 1) * step) - 1 : hi;
        lo = nlo;
        hi = max(nhi, nlo);
    }
    return lo;
}

}  // namespace raster
}  // namespace ipp
