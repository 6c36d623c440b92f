"""Plot calls are never reached (display_plot_frame=-1); any use is an error."""


def __getattr__(name):
    raise RuntimeError("matplotlib.pyplot.%s called: plots are out of scope" % name)
