/* Plain-C consumer of include/sfsim_atmosphere.h: proves the header is C (not C++), that the structs have
 * the documented layout and that the host-only entry points work without a GPU. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "sfsim_atmosphere.h"

int main(void) {
  atmlut_config cfg;
  memset(&cfg, 0, sizeof cfg);
  atmlut_default_config(&cfg);
  if (cfg.height_size != 32 || cfg.elevation_size != 127 || cfg.light_elevation_size != 32 || cfg.heading_size != 8 ||
      cfg.transmittance_height_size != 64 || cfg.transmittance_elevation_size != 255 || cfg.surface_height_size != 16 ||
      cfg.surface_sun_elevation_size != 63 || cfg.ray_steps != 100 || cfg.sphere_steps != 15 || cfg.iterations != 5)
    return 1;
  if (sizeof(atmlut_planet) != 64 || sizeof(atmlut_scatter) != 48 || offsetof(atmlut_config, intensity) != 48) return 2;
  int begin, count, per_rank;
  if (atmlut_slab(4064, 3, 8, &begin, &count, &per_rank) != 0 || begin != 1524 || count != 508 || per_rank != 508)
    return 3;
  if (atmlut_sphere_directions(15, 0, NULL, NULL, 0) != 71 || atmlut_sphere_directions(100, 1, NULL, NULL, 0) != 1605)
    return 4;
  float data[4] = {2.0f, 3.0f, 5.0f, 7.0f}, back[4] = {0, 0, 0, 0};
  if (atmlut_write_floats("c_abi_smoke.tmp", data, 4) != 0 || atmlut_read_floats("c_abi_smoke.tmp", back, 4) != 4 ||
      memcmp(data, back, sizeof data) != 0)
    return 5;
  remove("c_abi_smoke.tmp");
  printf("devices=%d last_error=\"%s\"\n", atmlut_device_count(), atmlut_last_error());
  return 0;
}
