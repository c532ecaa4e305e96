"""nafae_b200 -- B200-native (sm_100a) grounding hot path of jshi31/NAFAE.

Host side is plain Python/PyTorch; every operator calls hand-written CUDA kernels in
``libnafae_b200.so`` through its C ABI (``include/nafae_b200.h``).  There is no CPU path and no
fallback: importing an operator module (``grounding``, ``pipeline``, ``model.*``) without the built
library raises.  The pure-host modules (``synth``, ``evaluate``, ``checkpoint``, ``bridge`` and the
sharding helpers of ``parallel``) import without it -- the package itself loads nothing.

The sub-package ``nafae_b200.model`` mirrors the reference's ``lib/model`` layout, so reference code
written as ``from model.nms.nms_wrapper import nms`` keeps working with ``nafae_b200`` on ``sys.path``
(see INTEGRATION.md).
"""
__version__ = "0.1.0"

