"""mrcnn.graph of the reference (mrcnn/graph.py): the undirected graph whose depth-first connected components order the
merge groups of Analyzer.extract_det_masks.  The class lives in mrcnn.analyze (the batched path uses the host C++ routine
mrcnn_host_merge_components, which reproduces the same order); this module keeps `from mrcnn.graph import Graph` working."""
from .analyze import Graph

__all__ = ["Graph"]
