/* rt_oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE, see rt_oracle.c).
 * The same render signature is exported by the reference harness
 * (oracle/ref_harness.cpp -> ref_render) so tests can swap one for the other. */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/rt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Work counters of the reference algorithm.  A "ray" is one
 * calcIntersection call (raytracer.h:145): primary + shadow + refracted + reflected. */
typedef struct rt_oracle_counters {
  uint64_t rays;            /* calcIntersection calls                         */
  uint64_t shadow_rays;     /* of which hasClearLineOfSight                    */
  uint64_t sphere_tests;    /* raySphere calls = rays * N                      */
  uint64_t contain_queries; /* primaryContainer calls                          */
  uint64_t contain_tests;   /* primaryContainer loop iterations (early return) */
  uint64_t refractions;
  uint64_t reflections;
  uint64_t pops;
  uint64_t dropped_pushes;
  uint64_t samples;
  uint64_t max_stack;
} rt_oracle_counters;

/* Render rows row_begin + k*row_step, k in [0,row_count), of a W x H frame into
 * out[(k*W + x)*3 .. +2] (float RGB, un-normalised).  S = RTSTACK_MAXSIZE.
 * threads <= 0: all OpenMP threads.  Returns 0, or -1 on bad arguments. */
int rt_oracle_render(const rt_sphere* spheres, unsigned sphNum,
                     const rt_light* lights, unsigned lgtNum,
                     unsigned width, unsigned height, float zoom, float aliasFactor,
                     int maxStack, unsigned row_begin, unsigned row_count, unsigned row_step,
                     float* out, rt_oracle_counters* counters, int threads);

/* unit-level entry points for known-answer tests */
int rt_oracle_ray_sphere(const rt_sphere* s, const rt_vec* origin, const rt_vec* dir, float* t);
int rt_oracle_closest_hit(const rt_sphere* spheres, unsigned n, const rt_vec* origin,
                          const rt_vec* dir, rt_vec* point, rt_vec* normal, float* sqDist);
int rt_oracle_primary_container(const rt_sphere* spheres, unsigned n, const rt_vec* pt);
int rt_oracle_solve_quadratic(float a, float b, float c, float* roots);

float rt_oracle_max_colour(const float* rgb, size_t npixels);
void rt_oracle_quantise(const float* rgb, size_t npixels, float maxv, unsigned char* out);
int rt_oracle_threads(void);

#ifdef __cplusplus
}
#endif
#endif
