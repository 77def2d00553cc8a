"""hoomd_b200_plugin — a HOOMD-blue component (the layout of example_plugins/pair_plugin) that evaluates the LJ
pair force of an unmodified HOOMD-blue simulation with libhoomd_b200.so. Import `hoomd` first.

    import hoomd, hoomd_b200_plugin
    lj = hoomd_b200_plugin.pair.LJ(default_r_cut=2.5, buffer=0.4)
    lj.params[('A', 'A')] = dict(epsilon=1.0, sigma=1.0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[...], forces=[lj])

Built by baseline/build_plugin.sh against the reference's headers (git-ignored binary)."""
from . import pair  # noqa: F401
