/* rt_scene.h — scene construction for the trace loop (host side, plain C).
 *
 * The reference has no scene type (scene.h is empty; a scene is two arrays and
 * two counts, main.cpp:148-168) and hard-codes one scene in main().  These
 * helpers build that scene and the synthetic benchmark scenes of SURVEY.md
 * §8(d) into caller-owned arrays of the reference's own PODs.
 */
#ifndef RT_SCENE_H
#define RT_SCENE_H

#include <stdint.h>
#include "rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RT_SCENE_DEFAULT_SPHERES 3u
#define RT_SCENE_DEFAULT_LIGHTS  2u
#define RT_SCENE_MAX_SYNTH_LIGHTS 4u

/* Same result as setMatOpacity + setMatteGlossBalance + setMatRefractivityIndex
 * (raytracer.h:59-74): matte is pre-scaled by (float)(1.0 - glossFactor),
 * gloss by glossFactor. */
void rt_make_material(rt_material* out, const rt_vec* matte, const rt_vec* gloss,
                      float opacity, float glossFactor, float refractiveIndex);

/* The scene literal of main.cpp:113-168: 3 spheres, 2 lights. */
void rt_scene_default(rt_sphere spheres[RT_SCENE_DEFAULT_SPHERES],
                      rt_light lights[RT_SCENE_DEFAULT_LIGHTS]);

/* Deterministic synthetic scene synth(N, L, seed) of SURVEY.md §8(d): N random
 * spheres inside the reference camera's frustum, L <= 4 fixed point lights.
 * Returns 0, or -1 on bad arguments. */
int rt_scene_synth(unsigned sphNum, unsigned lgtNum, uint64_t seed,
                   rt_sphere* spheres, rt_light* lights);

/* Scene files (the reference has none: its scene is a literal in main()).  A text file,
 * one record per line, floats written as C99 hex-floats so a save/load round trip is exact:
 *     rtgamma-scene 1
 *     sphere  cx cy cz radius  matte.r matte.g matte.b  gloss.r gloss.g gloss.b  opacity refractiveIndex
 *     light   px py pz  col.r col.g col.b
 * Material fields are the stored ones (after rt_make_material); `#` starts a comment.
 * rt_scene_load allocates the two arrays with malloc (release with rt_scene_free).
 * Both return 0, or -1 on an I/O or syntax error. */
int rt_scene_save(const char* path, const rt_sphere* spheres, unsigned sphNum,
                  const rt_light* lights, unsigned lgtNum);
int rt_scene_load(const char* path, rt_sphere** spheres, unsigned* sphNum,
                  rt_light** lights, unsigned* lgtNum);
void rt_scene_free(void* arrayFromLoad);

#ifdef __cplusplus
}
#endif
#endif
