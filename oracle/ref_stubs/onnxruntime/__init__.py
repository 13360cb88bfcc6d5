"""Import stub so that the reference's infer/onnx.py (which holds the hot path's own schedule and
RoPE helpers) can be imported for fixture generation.  Sessions cannot be created."""


class SessionOptions:
    graph_optimization_level = None


class GraphOptimizationLevel:
    ORT_ENABLE_ALL = 99


def get_available_providers():
    return ["CPUExecutionProvider"]


class InferenceSession:
    def __init__(self, *a, **k):
        raise RuntimeError("onnxruntime is not available in this environment")
