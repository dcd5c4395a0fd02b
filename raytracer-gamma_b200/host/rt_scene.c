/* rt_scene.c — see include/rt_scene.h.  Plain C, no device code. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rt_scene.h"

static rt_vec vec3(float x, float y, float z) { rt_vec v = {x, y, z}; return v; }

void rt_make_material(rt_material* out, const rt_vec* matte, const rt_vec* gloss,
                      float opacity, float glossFactor, float refractiveIndex) {
  /* raytracer.h:67: the balance factor is formed in double, then narrowed by the
   * `float k = (a)` inside vsmul (vec.h:38) */
  const float km = (float)(1.0 - (double)glossFactor);
  const float kg = glossFactor;
  out->matteColour = vec3(km * matte->x, km * matte->y, km * matte->z);
  out->glossColour = vec3(kg * gloss->x, kg * gloss->y, kg * gloss->z);
  out->opacity = opacity;
  out->refractiveIndex = refractiveIndex;
}

void rt_scene_default(rt_sphere spheres[RT_SCENE_DEFAULT_SPHERES],
                      rt_light lights[RT_SCENE_DEFAULT_LIGHTS]) {
  /* colours, main.cpp:113-122 */
  const rt_vec lowerWhite = vec3(0.5f, 0.5f, 0.5f);
  const rt_vec redCol = vec3(0.8f, 1.f, 0.7f);
  const rt_vec greenCol = vec3(0.4f, 0.5f, 0.7f);
  const rt_vec col1 = vec3(0.01f, 0.8f, 0.01f);

  memset(spheres, 0, sizeof(rt_sphere) * RT_SCENE_DEFAULT_SPHERES);
  memset(lights, 0, sizeof(rt_light) * RT_SCENE_DEFAULT_LIGHTS);

  /* materials main.cpp:125-144, spheres :147-158 */
  rt_make_material(&spheres[0].material, &greenCol, &redCol, 0.8f, 0.2f, 1.5500f);
  spheres[0].pos = vec3(-9.f, 0.f, -13.f);  spheres[0].radius = 5.f;
  rt_make_material(&spheres[1].material, &greenCol, &redCol, 0.3f, 0.95f, 1.5500f);
  spheres[1].pos = vec3(-4.f, 1.5f, -5.f);  spheres[1].radius = 2.f;
  rt_make_material(&spheres[2].material, &col1, &col1, 0.6f, 0.0f, 1.5500f);
  spheres[2].pos = vec3(1.f, -1.f, -7.f);   spheres[2].radius = 3.f;

  /* lights main.cpp:161-168 */
  lights[0].pos = vec3(-45.f, 10.f, 85.f);  lights[0].col = lowerWhite;
  lights[1].pos = vec3(20.f, 60.f, -5.f);   lights[1].col = lowerWhite;
}

/* 64-bit LCG, 24-bit outputs */
static double lcg_next(uint64_t* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((*s >> 40) & 0xFFFFFFu) * (1.0 / 16777216.0);
}

int rt_scene_synth(unsigned sphNum, unsigned lgtNum, uint64_t seed,
                   rt_sphere* spheres, rt_light* lights) {
  static const float lightPos[RT_SCENE_MAX_SYNTH_LIGHTS][3] = {
    {-45.f, 10.f, 85.f}, {20.f, 60.f, -5.f}, {40.f, -30.f, 60.f}, {-20.f, 80.f, 20.f}};
  if (sphNum == 0 || lgtNum > RT_SCENE_MAX_SYNTH_LIGHTS || !spheres || (lgtNum && !lights))
    return -1;
  uint64_t s = 0x9E3779B97F4A7C15ULL ^ (uint64_t)sphNum ^ seed;
  const double shrink = cbrt(256.0 / (double)sphNum);
  for (unsigned i = 0; i < sphNum; ++i) {
    const double z = -24.0 + 18.0 * lcg_next(&s);
    const double az = fabs(z);
    const double x = (2.0 * lcg_next(&s) - 1.0) * 0.9 * 2.667 * az;
    const double y = (2.0 * lcg_next(&s) - 1.0) * 0.9 * 1.5 * az;
    const double r = (0.25 + 0.75 * lcg_next(&s)) * shrink * az / 10.0;
    rt_vec matte, gloss;
    matte.x = (float)(0.1 + 0.9 * lcg_next(&s));
    matte.y = (float)(0.1 + 0.9 * lcg_next(&s));
    matte.z = (float)(0.1 + 0.9 * lcg_next(&s));
    gloss.x = (float)(0.1 + 0.9 * lcg_next(&s));
    gloss.y = (float)(0.1 + 0.9 * lcg_next(&s));
    gloss.z = (float)(0.1 + 0.9 * lcg_next(&s));
    const double pick = lcg_next(&s);
    const double op = 0.3 + 0.7 * lcg_next(&s);
    const float opacity = (pick < 0.25) ? 1.f : (float)op;
    const float glossFactor = (float)(0.9 * lcg_next(&s));
    const float index = (float)(1.1 + 0.7 * lcg_next(&s));
    memset(&spheres[i], 0, sizeof spheres[i]);
    spheres[i].pos = vec3((float)x, (float)y, (float)z);
    spheres[i].radius = (float)r;
    rt_make_material(&spheres[i].material, &matte, &gloss, opacity, glossFactor, index);
  }
  for (unsigned l = 0; l < lgtNum; ++l) {
    lights[l].pos = vec3(lightPos[l][0], lightPos[l][1], lightPos[l][2]);
    lights[l].col = vec3(0.5f, 0.5f, 0.5f);
  }
  return 0;
}

int rt_scene_save(const char* path, const rt_sphere* spheres, unsigned sphNum,
                  const rt_light* lights, unsigned lgtNum) {
  if (!path || (sphNum && !spheres) || (lgtNum && !lights)) return -1;
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "rtgamma-scene 1\n# %u spheres, %u lights\n", sphNum, lgtNum);
  for (unsigned i = 0; i < sphNum; ++i) {
    const rt_sphere* s = &spheres[i];
    fprintf(f, "sphere %a %a %a %a  %a %a %a  %a %a %a  %a %a\n", s->pos.x, s->pos.y, s->pos.z, s->radius,
            s->material.matteColour.x, s->material.matteColour.y, s->material.matteColour.z,
            s->material.glossColour.x, s->material.glossColour.y, s->material.glossColour.z,
            s->material.opacity, s->material.refractiveIndex);
  }
  for (unsigned l = 0; l < lgtNum; ++l)
    fprintf(f, "light %a %a %a  %a %a %a\n", lights[l].pos.x, lights[l].pos.y, lights[l].pos.z,
            lights[l].col.x, lights[l].col.y, lights[l].col.z);
  const int bad = ferror(f);
  return (fclose(f) == 0 && !bad) ? 0 : -1;
}

void rt_scene_free(void* p) { free(p); }

int rt_scene_load(const char* path, rt_sphere** spheres, unsigned* sphNum,
                  rt_light** lights, unsigned* lgtNum) {
  if (!path || !spheres || !sphNum || !lights || !lgtNum) return -1;
  *spheres = NULL; *lights = NULL; *sphNum = 0; *lgtNum = 0;
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  char line[1024];
  unsigned ns = 0, nl = 0, capS = 0, capL = 0;
  rt_sphere* S = NULL; rt_light* L = NULL;
  int ok = 1, header = 0;
  while (ok && fgets(line, sizeof line, f)) {
    char* hash = strchr(line, '#');
    if (hash) *hash = 0;
    char tag[32];
    if (sscanf(line, "%31s", tag) != 1) continue;           /* blank line */
    if (!header) { int v = 0; ok = (!strcmp(tag, "rtgamma-scene") && sscanf(line, "%*s %d", &v) == 1 && v == 1); header = 1; continue; }
    if (!strcmp(tag, "sphere")) {
      float v[12];
      if (sscanf(line, "%*s %f %f %f %f %f %f %f %f %f %f %f %f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6],
                 &v[7], &v[8], &v[9], &v[10], &v[11]) != 12) { ok = 0; break; }
      if (ns == capS) { capS = capS ? capS * 2 : 64; rt_sphere* t = (rt_sphere*)realloc(S, capS * sizeof *t); if (!t) { ok = 0; break; } S = t; }
      memcpy(&S[ns++], v, sizeof(rt_sphere));               /* rt_sphere is 12 packed floats in this order */
    } else if (!strcmp(tag, "light")) {
      float v[6];
      if (sscanf(line, "%*s %f %f %f %f %f %f", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) != 6) { ok = 0; break; }
      if (nl == capL) { capL = capL ? capL * 2 : 8; rt_light* t = (rt_light*)realloc(L, capL * sizeof *t); if (!t) { ok = 0; break; } L = t; }
      memcpy(&L[nl++], v, sizeof(rt_light));
    } else {
      ok = 0;
    }
  }
  fclose(f);
  if (!ok || !header) { free(S); free(L); return -1; }
  *spheres = S; *sphNum = ns; *lights = L; *lgtNum = nl;
  return 0;
}
