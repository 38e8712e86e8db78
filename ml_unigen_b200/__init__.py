"""unigen-b200: B200-native (sm_100a) text-to-image parallel decoding for apple/ml-unigen.

Host-side mirror of the reference's interface for this path (same names, arguments and error
behaviour), calling hand-written CUDA through the C ABI in include/unigen_b200.h:

    ml_unigen_b200.sampling   <- models/sampling.py      (mask_by_random_topk, schedules)
    ml_unigen_b200.prompting  <- training/prompting_utils.py:113-153, 975-1020
    ml_unigen_b200.unigen     <- models/unigen.py:344-455 (UniGen.t2i_generate)
    ml_unigen_b200.dist       <- evaluation/inference_unigen_cot.py:146-243, 417-427 (shard + gather)
    ml_unigen_b200.magvit     <- models/multimodal_encoder/magvitv2.py:217-230, 286-408, 438-442 (decode_code)
    ml_unigen_b200.callers    <- evaluation/inference_geneval.py:152-203, 266-276; inference_unigen_cot.py:507-534
                                 (prompt files, per-rank prompts, output layout, CoT-V top-k selection)

There is no CPU fallback: importing the kernels without the built library raises.
"""
__version__ = "0.1.0"
