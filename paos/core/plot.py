"""``paos.core.plot`` (reference ``paos/core/plot.py``): plotting is outside the device path (matplotlib is not part of
this image); the name exists so that ``import paos`` matches the reference, and calling it says where to go."""


def plot_pop(*args, **kwargs):
    raise NotImplementedError("plotting is not part of paos_b200: hand the dictionaries returned by run / pipeline to the "
                              "reference's paos.core.plot.plot_pop")


simple_plot = plot_psf_xsec = plot_surface = plot_pop
