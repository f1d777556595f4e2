"""sfsim_b200 -- host-side layer of the B200 atmosphere-LUT accelerator for sfsim.

Everything computes on the GPU through libsfsim_atmosphere.so (include/sfsim_atmosphere.h); there is no
CPU fallback.  Modules mirror the reference's namespaces:

    sfsim_b200.atmosphere_lut   sfsim.atmosphere-lut   generate_atmosphere_luts, AtmosphereLutBuilder
    sfsim_b200.atmosphere       sfsim.atmosphere       transmittance, ray_scatter, point_scatter, index maps, spaces
    sfsim_b200.interpolate      sfsim.interpolate      make_lookup_table, interpolation_table, linear_space, ...
    sfsim_b200.sharding                                slab arithmetic of the multi-GPU build
"""
