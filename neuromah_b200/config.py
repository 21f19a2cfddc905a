"""Run-time switches for the reference's quirks.

Defaults reproduce the reference's numerics on the north-star hot path:
  Q1  every trainable layer registered twice by Model.finalize -> optimizer applied twice
  Q2  Dense gradients divided by the batch a second time        (always on; it is Dense.py's math)
  Q3  Layer_Conv2D.backward: "wired" = the intended dgrad/wgrad (north star), "stub" = as shipped
  Q4  Layer_Conv2D.forward does not add its bias
"""

duplicate_trainable_registration = True     # Q1 (Model.py:56-59 and :82-83)
conv_backward = "wired"                     # Q3: "wired" | "stub" (Conv_wrapepr.py:100-115)
conv_bias_in_forward = False                # Q4 (Conv_wrapepr.py:86-98)
