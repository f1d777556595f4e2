// Double-precision geometry, medium and index maps of sfsim.atmosphere, as host/device inline
// functions.  These define every texel's inputs and every lookup's address, so they are evaluated
// exactly as the reference writes them (operation order kept; the translation unit is compiled
// with -fmad=false so nothing is contracted behind our back).
//
// file:line citations refer to wedesoft/sfsim (src/clj/sfsim/...).
#pragma once

#include <cmath>
#include <cuda_runtime.h>

#define HD __host__ __device__ __forceinline__

namespace atm {

constexpr double kPi = 3.14159265358979323846;

struct V3 {
  double x, y, z;
};

HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
HD V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HD double mag(V3 a) { return sqrt(dot(a, a)); }
HD double sqr(double x) { return x * x; }  // util.clj:334-337

// planet (atmosphere.clj:64-67) with the centre already subtracted from every point
struct Planet {
  double radius, height;
  double brightness[3];
};

// scatter components (atmosphere.clj:35-39); n = 1 or 2 like the reference's transmittance
struct Medium {
  int n;
  double base[2][3];
  double scale[2];
  double g[2];
  double quotient[2];
};

// util.clj:403-416 limit-quot (the recursion for b < 0 unrolled)
HD double limit_quot(double a, double b, double lower, double upper) {
  if (a == 0.0) return a;
  if (b < 0) {
    a = -a;
    b = -b;
  }
  if (a < b * upper) {
    if (a > b * lower) return a / b;
    return lower;
  }
  return upper;
}

// sphere.clj:28-31 height (centre = origin)
HD double height(const Planet &p, V3 point) { return mag(point) - p.radius; }

// sphere.clj:34-59 ray-sphere-intersection (centre = origin)
HD void ray_sphere_intersection(double radius, V3 origin, V3 direction, double &distance, double &length) {
  double direction_sqr = dot(direction, direction);
  double discriminant = sqr(dot(direction, origin)) - direction_sqr * (dot(origin, origin) - sqr(radius));
  double middle = -(dot(direction, origin) / direction_sqr);
  if (discriminant > 0) {
    double length2 = sqrt(discriminant) / direction_sqr;
    if (middle < length2) {
      distance = 0.0;
      length = fmax(0.0, middle + length2);
    } else {
      distance = middle - length2;
      length = 2 * length2;
    }
  } else {
    distance = fmax(0.0, middle);
    length = 0.0;
  }
}

// atmosphere.clj:70-77 atmosphere-intersection
HD V3 atmosphere_intersection(const Planet &p, V3 origin, V3 direction) {
  double distance, length;
  ray_sphere_intersection(p.radius + p.height, origin, direction, distance, length);
  return origin + direction * (distance + length);
}

// atmosphere.clj:80-85 surface-intersection
HD V3 surface_intersection(const Planet &p, V3 origin, V3 direction) {
  double distance, length;
  ray_sphere_intersection(p.radius, origin, direction, distance, length);
  return origin + direction * distance;
}

// atmosphere.clj:88-92 surface-point?
HD bool surface_point(const Planet &p, V3 point) { return 2.0 * height(p, point) < p.height; }

// atmosphere.clj:95-102 is-above-horizon?
HD bool is_above_horizon(const Planet &p, V3 point, V3 direction) {
  double norm_point = mag(point);
  double sin_elevation_radius = dot(direction, point);
  double horizon_distance_sqr = sqr(norm_point) - sqr(p.radius);
  return sin_elevation_radius >= 0 || sqr(sin_elevation_radius) <= horizon_distance_sqr;
}

// atmosphere.clj:105-111 ray-extremity
HD V3 ray_extremity(const Planet &p, V3 origin, V3 direction) {
  return is_above_horizon(p, origin, direction) ? atmosphere_intersection(p, origin, direction)
                                                : surface_intersection(p, origin, direction);
}

// atmosphere.clj:56-61 phase
HD double phase(double g, double mu) {
  double g2 = sqr(g);
  return (3.0 * (1.0 - g2) * (1.0 + sqr(mu))) / (8.0 * kPi * (2.0 + g2) * pow((1.0 + g2) - 2.0 * g * mu, 1.5));
}

// atmosphere.clj:42-47 scattering of component c, one colour channel
HD double scattering(const Medium &m, int c, int ch, double h) { return m.base[c][ch] * exp(-(h / m.scale[c])); }

// atmosphere.clj:118-125 transmittance between two points, plain double-precision restatement
// (integral-ray, ray.clj:19-30, with overall-extinction atmosphere.clj:119-124)
HD void transmittance_points(const Planet &p, const Medium &m, int steps, V3 x, V3 x0, double out[3]) {
  V3 d = x0 - x;
  double stepsize = 1.0 / (double)steps;
  double a = stepsize * mag(d);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int n = 0; n < steps; n++) {
    double s = (0.5 + (double)n) * stepsize;
    V3 q = x + d * s;
    double h = height(p, q);
    double e[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < m.n; c++) {
      double dens = exp(-(h / m.scale[c]));
      for (int ch = 0; ch < 3; ch++) {
        double ext = (m.base[c][ch] * dens) / m.quotient[c];
        e[ch] = (c == 0) ? ext : e[ch] + ext;
      }
    }
    for (int ch = 0; ch < 3; ch++) acc[ch] = acc[ch] + e[ch] * a;
  }
  for (int ch = 0; ch < 3; ch++) out[ch] = exp(-acc[ch]);
}

// atmosphere.clj:126-128 transmittance towards the shell or the ground
HD void transmittance_dir(const Planet &p, const Medium &m, int steps, V3 x, V3 v, bool above, double out[3]) {
  V3 x0 = above ? atmosphere_intersection(p, x, v) : surface_intersection(p, x, v);
  transmittance_points(p, m, steps, x, x0, out);
}

// ------------------------------------------------------------------ index maps, atmosphere.clj:233-422

// atmosphere.clj:233-236
HD double horizon_distance(const Planet &p, double radius) { return sqrt(fmax(0.0, sqr(radius) - sqr(p.radius))); }

// atmosphere.clj:239-253
HD double elevation_to_index(const Planet &p, int size, V3 point, V3 direction, bool above) {
  double radius = mag(point);
  double ground_radius = p.radius;
  double top_radius = ground_radius + p.height;
  double sin_elevation = dot(point, direction) / radius;
  double rho = horizon_distance(p, radius);
  double Delta = sqr(radius * sin_elevation) - sqr(rho);
  double H = sqrt(sqr(top_radius) - sqr(ground_radius));
  double q;
  if (above)
    q = 0.5 - limit_quot(radius * sin_elevation - sqrt(fmax(0.0, Delta + sqr(H))), rho + rho + 2 * H, -0.5, 0.0);
  else
    q = 0.5 + limit_quot(radius * sin_elevation + sqrt(fmax(0.0, Delta)), rho + rho, -0.5, 0.0);
  return (double)(size - 1) * q;
}

// atmosphere.clj:256-270
HD void index_to_elevation(const Planet &p, int size, double radius, double index, V3 &dir, bool &above) {
  double ground_radius = p.radius;
  double top_radius = ground_radius + p.height;
  double horizon_dist = horizon_distance(p, radius);
  double H = sqrt(sqr(top_radius) - sqr(ground_radius));
  double scaled_index = index / (double)(size - 1);
  double sin_elevation;
  if (scaled_index < 0.5 || (scaled_index == 0.5 && 2.0 * radius < ground_radius + top_radius)) {
    double ground_dist = horizon_dist * (1 - 2 * scaled_index);
    sin_elevation =
        limit_quot(sqr(ground_radius) - sqr(radius) - sqr(ground_dist), 2 * radius * ground_dist, -1.0, 1.0);
    above = false;
  } else {
    double sky_dist = (horizon_dist + H) * (2 * scaled_index - 1);
    sin_elevation = limit_quot(sqr(top_radius) - sqr(radius) - sqr(sky_dist), 2 * radius * sky_dist, -1.0, 1.0);
    above = true;
  }
  dir = V3{sin_elevation, sqrt(1 - sqr(sin_elevation)), 0.0};
}

// atmosphere.clj:273-278
HD double height_to_index(const Planet &p, int size, V3 point) {
  return (double)(size - 1) * (horizon_distance(p, mag(point)) / horizon_distance(p, p.radius + p.height));
}

// atmosphere.clj:281-288
HD V3 index_to_height(const Planet &p, int size, double index) {
  double max_horizon = sqrt(sqr(p.radius + p.height) - sqr(p.radius));
  double horizon_dist = (index / (double)(size - 1)) * max_horizon;
  return V3{sqrt(sqr(p.radius) + sqr(horizon_dist)), 0.0, 0.0};
}

// atmosphere.clj:322-326, from the sine of the sun elevation
HD double sin_sun_elevation_to_index(int size, double sin_elevation) {
  return (double)(size - 1) * fmax(0.0, (1 - exp(0 - 3 * sin_elevation - 0.6)) / (1 - exp(-3.6)));
}
// Same map for the lookups inside the integration kernels: the division by the constant 1 - exp(-3.6) is a
// multiplication by its reciprocal (the result is continuous in the coordinate and the clamp to exactly 0
// happens before the scaling, so the one-ulp difference cannot change a table value beyond rounding).
HD double sin_sun_elevation_to_index_fast(int size, double sin_elevation) {
  const double inv = 1.0 / (1 - 0.02732372244729256);   // 1 / (1 - exp(-3.6))
  return (double)(size - 1) * fmax(0.0, (1 - exp(0 - 3 * sin_elevation - 0.6)) * inv);
}
HD double sun_elevation_to_index(int size, V3 point, V3 light) {
  return sin_sun_elevation_to_index(size, dot(point, light) / mag(point));
}

// atmosphere.clj:329-332
HD double index_to_sin_sun_elevation(int size, double index) {
  return (log(1 - (index / (double)(size - 1)) * (1 - exp(-3.6))) + 0.6) / -3;
}

// atmosphere.clj:368-372
HD double sun_angle_to_index(int size, V3 direction, V3 light) {
  return (double)(size - 1) * ((1 + dot(direction, light)) / 2);
}

// atmosphere.clj:375-384
HD V3 index_to_sun_direction(int size, V3 direction, double sin_sun_elevation, double index) {
  double dot_view_sun = 2.0 * (index / (double)(size - 1)) - 1.0;
  double max_sun_1 = sqrt(fmax(0.0, 1.0 - sqr(sin_sun_elevation)));
  double sun_1 = limit_quot(dot_view_sun - direction.x * sin_sun_elevation, direction.y, -max_sun_1, max_sun_1);
  double sun_2 = sqrt(fmax(0.0, 1.0 - sqr(sun_1) - sqr(sin_sun_elevation)));
  return V3{sin_sun_elevation, sun_1, sun_2};
}

// atmosphere.clj:401-412 ray-scatter-backward
HD void ray_scatter_backward(const Planet &p, const int shape[4], double hi, double ei, double si, double ai, V3 &point,
                             V3 &direction, V3 &light, bool &above) {
  point = index_to_height(p, shape[0], hi);
  index_to_elevation(p, shape[1], point.x, ei, direction, above);
  double sin_sun_elevation = index_to_sin_sun_elevation(shape[2], si);
  light = index_to_sun_direction(shape[3], direction, sin_sun_elevation, ai);
}

// atmosphere.clj:387-398 ray-scatter-forward
HD void ray_scatter_forward(const Planet &p, const int shape[4], V3 point, V3 direction, V3 light, bool above,
                            double idx[4]) {
  idx[0] = height_to_index(p, shape[0], point);
  idx[1] = elevation_to_index(p, shape[1], point, direction, above);
  idx[2] = sun_elevation_to_index(shape[2], point, light);
  idx[3] = sun_angle_to_index(shape[3], direction, light);
}

// atmosphere.clj:346-356 surface-radiance-backward
HD void surface_radiance_backward(const Planet &p, const int shape[2], double hi, double si, V3 &point, V3 &light) {
  point = index_to_height(p, shape[0], hi);
  double sin_sun_elevation = index_to_sin_sun_elevation(shape[1], si);
  double cos_sun_elevation = sqrt(fmax(0.0, 1 - sqr(sin_sun_elevation)));
  light = V3{sin_sun_elevation, cos_sun_elevation, 0.0};
}

// atmosphere.clj:335-343 surface-radiance-forward
HD void surface_radiance_forward(const Planet &p, const int shape[2], V3 point, V3 light, double idx[2]) {
  idx[0] = height_to_index(p, shape[0], point);
  idx[1] = sun_elevation_to_index(shape[1], point, light);
}

// atmosphere.clj:302-310 transmittance-backward
HD void transmittance_backward(const Planet &p, const int shape[2], double hi, double ei, V3 &point, V3 &direction,
                               bool &above) {
  point = index_to_height(p, shape[0], hi);
  index_to_elevation(p, shape[1], point.x, ei, direction, above);
}

// atmosphere.clj:291-299 transmittance-forward
HD void transmittance_forward(const Planet &p, const int shape[2], V3 point, V3 direction, bool above,
                              double idx[2]) {
  idx[0] = height_to_index(p, shape[0], point);
  idx[1] = elevation_to_index(p, shape[1], point, direction, above);
}

}  // namespace atm
