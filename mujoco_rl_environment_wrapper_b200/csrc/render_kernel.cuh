// render_kernel.cuh — agent cameras (reference: MuJoCo_Gym/mujoco_parent.py:496-575,
// `get_camera_data` -> mjv_updateScene + mjr_render + mjr_readPixels on an off-screen GLFW window).
//
// The reference rasterises with MuJoCo's OpenGL renderer; here every pixel is a primary ray against the
// env's primitive geoms (the same ray / geom routines the rangefinder uses), shaded with a two-sided
// head-light:  rgb = clamp(geom_rgba.rgb, 0, 1) * (RENDER_AMBIENT + RENDER_DIFFUSE * |n . d|), black where
// nothing is hit, 8 bits per channel, rows bottom-up like glReadPixels.  Lights, materials, textures,
// shadows and fog of the OpenGL pipeline are not modelled (pixel parity with the reference is not claimed;
// the CPU oracle restates exactly this image formation in fp64).
//
// One CTA per env: warp 0 redoes the kinematics pass from the env's qpos, the CTA flattens the geoms into a
// broadcast-friendly list in shared memory, then every thread shades groups of 4 neighbouring pixels and
// writes them as three 32-bit words (coalesced 384 B per warp).
#pragma once
#include "step_kernel.cuh"

namespace mjb {

#define RENDER_AMBIENT 0.4f
#define RENDER_DIFFUSE 0.6f
#define RENDER_MAX_CAMS 8

struct RenderCams { int n; int id[RENDER_MAX_CAMS]; };

// colour of the first geom along the ray (o, unit d); packed 0x00BBGGRR
MJB_DEV uint32_t shade_ray(const float* GL, int ngeom, f3 o, f3 d) {
  float best = MJB_BIG;
  int bi = -1;
  MJB_NOUNROLL
  for (int g = 0; g < ngeom; g++) {
    const float* G = GL + g * GL_STRIDE;
    const int type = __float_as_int(G[GL_TYPE]);
    if (type < 0) continue;  // invisible (alpha == 0)
    f3 pos = ld3(G + GL_POS);
    if (type != MJB_GEOM_PLANE) {  // bounding sphere first
      f3 oc = pos - o;
      float tca = dot(oc, d), l2 = dot(oc, oc), rb = G[GL_RBOUND];
      if (l2 - tca * tca > rb * rb || (tca < 0.f && l2 > rb * rb) || tca - rb > best) continue;
    }
    float t = ray_geom(pos, G + GL_MAT, G + GL_SIZE, o, d, type);
    if (t >= 0.f && t < best) { best = t; bi = g; }
  }
  if (bi < 0) return 0u;
  const float* G = GL + bi * GL_STRIDE;
  const int type = __float_as_int(G[GL_TYPE]);
  const float* R = G + GL_MAT;
  float nd;  // |normal . direction|
  if (type == MJB_GEOM_PLANE) {
    nd = fabsf(dot(colv(R, 2), d));
  } else if (type == MJB_GEOM_SPHERE) {
    f3 n = (o + d * best) - ld3(G + GL_POS);
    nd = fabsf(dot(n, d)) * MJB_RSQRT(fmaxf(dot(n, n), 1e-20f));
  } else {
    f3 lp = mulTv(R, (o + d * best) - ld3(G + GL_POS)), ld = mulTv(R, d);
    if (type == MJB_GEOM_CAPSULE) {
      float h = G[GL_SIZE + 1];
      f3 n = mk3(lp.x, lp.y, lp.z - fminf(fmaxf(lp.z, -h), h));
      nd = fabsf(dot(n, ld)) * MJB_RSQRT(fmaxf(dot(n, n), 1e-20f));
    } else {  // box: the face whose scaled coordinate is largest
      float ax = fabsf(lp.x) / G[GL_SIZE], ay = fabsf(lp.y) / G[GL_SIZE + 1], az = fabsf(lp.z) / G[GL_SIZE + 2];
      nd = (ax >= ay && ax >= az) ? fabsf(ld.x) : (ay >= az ? fabsf(ld.y) : fabsf(ld.z));
    }
  }
  const float I = RENDER_AMBIENT + RENDER_DIFFUSE * fminf(nd, 1.f);
  uint32_t r = (uint32_t)(G[GL_RGB] * I * 255.f + 0.5f), gg = (uint32_t)(G[GL_RGB + 1] * I * 255.f + 0.5f),
           bb = (uint32_t)(G[GL_RGB + 2] * I * 255.f + 0.5f);
  return r | (gg << 8) | (bb << 16);
}

#if !defined(MJB_HOST_EMU)
// out: u8 [num_envs, cams.n, height, width, 3]
__global__ void __launch_bounds__(256) k_render(const __grid_constant__ DevModel dm, const uint32_t* __restrict__ image,
                                                const RenderHdr rh, const uint32_t* __restrict__ rtab,
                                                const float* __restrict__ qpos_g, int num_envs, const RenderCams cams, int width,
                                                int height, uint8_t* __restrict__ out) {
  extern __shared__ __align__(128) uint32_t smem[];
  uint32_t* img = smem;
  float* scratch = reinterpret_cast<float*>(img + dm.image_words);
  uint32_t* tab = reinterpret_cast<uint32_t*>(scratch + dm.env_words);
  float* GL = reinterpret_cast<float*>(tab + rh.words);
  float* CW = GL + dm.ngeom * GL_STRIDE;   // per requested camera: world pos[3], rotation[9]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < dm.image_words; i += blockDim.x) img[i] = image[i];
  for (int i = tid; i < rh.words; i += blockDim.x) tab[i] = rtab[i];
  __syncthreads();
  Ctx c{&dm, img, scratch, lane, nullptr, 0, 0};
  const float* rgba = reinterpret_cast<const float*>(tab + rh.off_rgba);
  const int px = (width & 3) ? 1 : 4;            // pixels per work item
  const int items_row = width / px, items_cam = items_row * height, items = items_cam * cams.n;
  for (int env = blockIdx.x; env < num_envs; env += gridDim.x) {
    if (tid < 32) {
      float* qpos = SF(qpos);
      for (int i = lane; i < dm.nq; i += 32) qpos[i] = qpos_g[(size_t)env * dm.qpos_stride + i];
      __syncwarp();
      fk(c);
    }
    __syncthreads();
    for (int g = tid; g < dm.ngeom; g += blockDim.x) {
      GeomW w = geom_world(c, g);
      float* G = GL + g * GL_STRIDE;
      st3(G + GL_POS, w.pos);
      for (int i = 0; i < 9; i++) G[GL_MAT + i] = w.mat[i];
      for (int i = 0; i < 3; i++) { G[GL_SIZE + i] = w.size[i]; G[GL_RGB + i] = rgba[4 * g + i]; }
      G[GL_TYPE] = __int_as_float(rgba[4 * g + 3] > 0.f ? w.type : -1);
      G[GL_RBOUND] = CF(geom_rbound)[g];
    }
    for (int k = tid; k < cams.n; k += blockDim.x) {
      const uint32_t* cr = tab + rh.off_cam + cams.id[k] * CAM_STRIDE;
      const float* cf = reinterpret_cast<const float*>(cr);
      int mb = (int)cr[CAM_MB];
      f3 p = ld3(cf + CAM_POS);
      q4 q = ldq(cf + CAM_QUAT);
      if (mb >= 0) { p = ld3(SF(xpos) + 3 * mb) + mulv(SF(xmat) + 9 * mb, p); q = qmul(ldq(SF(xquat) + 4 * mb), q); }
      st3(CW + 12 * k, p);
      q2m(q, CW + 12 * k + 3);
    }
    __syncthreads();
    for (int it = tid; it < items; it += blockDim.x) {
      const int k = it / items_cam, r = it - k * items_cam, iy = r / items_row, ix0 = (r - iy * items_row) * px;
      const float* cw = CW + 12 * k;
      const float th = reinterpret_cast<const float*>(tab + rh.off_cam + cams.id[k] * CAM_STRIDE)[CAM_TANHALF];
      const f3 o = ld3(cw);
      const float v = ((iy + 0.5f) / height * 2.f - 1.f) * th;
      uint32_t col[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (j >= px) break;
        float u = ((ix0 + j + 0.5f) / width * 2.f - 1.f) * th * ((float)width / (float)height);
        f3 dl = mk3(u, v, -1.f);   // camera frame: looks along -z, +y up
        f3 d = mulv(cw + 3, dl * MJB_RSQRT(dot(dl, dl)));
        col[j] = shade_ray(GL, dm.ngeom, o, d);
      }
      uint8_t* dst = out + ((((size_t)env * cams.n + k) * height + iy) * width + ix0) * 3;
      if (px == 4) {  // 12 bytes = three aligned words
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
        d32[0] = col[0] | (col[1] << 24);
        d32[1] = (col[1] >> 8) | (col[2] << 16);
        d32[2] = (col[2] >> 16) | (col[3] << 8);
      } else {
        dst[0] = col[0] & 0xff; dst[1] = (col[0] >> 8) & 0xff; dst[2] = (col[0] >> 16) & 0xff;
      }
    }
    __syncthreads();   // the scratch is reused by the next env
  }
}
#endif

}  // namespace mjb
