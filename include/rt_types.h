/* rt_types.h — plain-float scene PODs shared by the C-ABI, the host and the oracle.
 *
 * Layout-identical to the reference's live structs (all float, 4-byte aligned,
 * no padding), so a `struct Sphere*` / `struct Light*` / `Vec*` from the
 * reference can be passed straight through the C-ABI:
 *   rt_vec      <-> Vec              vec.h:27-29          12 B
 *   rt_material <-> struct Material  material.h:8-14      32 B
 *   rt_sphere   <-> struct Sphere    sphere.h:9-14        48 B
 *   rt_light    <-> struct Light     raytracer.h:20-25    24 B  (NOT the stale light.h)
 *   rt_ray      <-> struct Ray       ray.h:9-14           36 B
 */
#ifndef RT_TYPES_H
#define RT_TYPES_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_vec { float x, y, z; } rt_vec;

typedef struct rt_material {
  rt_vec matteColour;
  rt_vec glossColour;
  float opacity;
  float refractiveIndex;
} rt_material;

typedef struct rt_sphere {
  rt_vec pos;
  float radius;
  rt_material material;
} rt_sphere;

typedef struct rt_light {
  rt_vec pos;
  rt_vec col;
} rt_light;

typedef struct rt_ray {
  rt_vec origin;
  rt_vec dir;
  rt_vec intensity;
} rt_ray;

#ifdef __cplusplus
}
static_assert(sizeof(rt_vec) == 12, "Vec ABI");
static_assert(sizeof(rt_material) == 32, "Material ABI");
static_assert(sizeof(rt_sphere) == 48, "Sphere ABI");
static_assert(sizeof(rt_light) == 24, "Light ABI");
static_assert(sizeof(rt_ray) == 36, "Ray ABI");
#else
_Static_assert(sizeof(rt_vec) == 12, "Vec ABI");
_Static_assert(sizeof(rt_material) == 32, "Material ABI");
_Static_assert(sizeof(rt_sphere) == 48, "Sphere ABI");
_Static_assert(sizeof(rt_light) == 24, "Light ABI");
_Static_assert(sizeof(rt_ray) == 36, "Ray ABI");
#endif

#endif /* RT_TYPES_H */
