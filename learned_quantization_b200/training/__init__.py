"""Host-side training-step plumbing around the hot path:

    dp.py      one NCCL communicator per process through the C-ABI (lq_dp_*), flat gradient bucket
    step.py    QuantBank (all quantised tensors of a model in ONE launch per phase),
               TrainStep (forward, backward, gradient exchange, fused optimizer + constraint)
    models.py  the reference's three model builders on the drop-in layers (benchmark harness)

The reference has no counterpart for dp.py / step.py: it is a single-device Keras `model.fit`
(IMAGENETTE/nested_quantization_layer/experiment.py:774-788).
"""
from .dp import Comm, NcclComm, P2PFinish, TorchDistComm, GradBucket, exchange_gradients  # noqa: F401
