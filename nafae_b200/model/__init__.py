"""Mirror of the reference's ``lib/model`` package layout for the hot-path operators.

Reference import                                         -> provided here
``from model.nms.nms_wrapper import nms``                -> model/nms/nms_wrapper.py
``from model.roi_align.modules.roi_align import ...``    -> model/roi_align/modules/roi_align.py
``from model.roi_pooling.modules.roi_pool import ...``   -> model/roi_pooling/modules/roi_pool.py
(new) ``model.rpn.proposal_layer.proposal_tail``         -> batched form of proposal_layer.py:127-163
"""
