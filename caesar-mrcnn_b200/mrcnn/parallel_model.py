"""mrcnn.parallel_model of the reference (mrcnn/parallel_model.py:22-104): Keras towers on every GPU of ONE process, outputs
merged on the CPU.  The B200 build scales with one process per GPU instead (torchrun; gradient all-reduce in
mrcnn.training.GradReducer, image / tile sharding for detection), so there is nothing to wrap: the name is kept so that
imports of the reference's scripts resolve, and using it says what to do instead."""


class ParallelModel(object):
    def __init__(self, keras_model=None, gpu_count=1):
        raise NotImplementedError("ParallelModel (multi-GPU towers inside one process) is replaced by one process per GPU: "
                                  "launch with `torchrun --nproc-per-node N ...` and use MaskRCNN directly (DESIGN.md §5, §11)")
