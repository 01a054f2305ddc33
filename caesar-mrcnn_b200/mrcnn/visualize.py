"""mrcnn.visualize of the reference (matplotlib / IPython plotting helpers, mrcnn/visualize.py:35-500): outside the rebuilt
detect / train path (out of scope; matplotlib is not part of this image).  The module exists so that
`from mrcnn import visualize` at the top of the reference's scripts resolves; every plotting entry point raises."""

_NAMES = ("display_images", "random_colors", "apply_mask", "display_instances", "display_differences", "draw_rois", "draw_box",
          "display_top_masks", "plot_precision_recall", "plot_overlaps", "draw_boxes", "display_table", "display_weight_stats")


def _unavailable(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("mrcnn.visualize.%s: plotting is not provided by the B200 build (DESIGN.md §6)" % name)
    fn.__name__ = name
    return fn


for _n in _NAMES:
    globals()[_n] = _unavailable(_n)
del _n
