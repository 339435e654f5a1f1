"""scanpaths_b200: B200-native (sm_100a) scanpath decoding, sampling and scoring.

Drop-in for the hot path of chenxy99/Scanpaths; the mirror modules keep the
reference's import names:
    scanpaths_b200.models.sampling.Sampling
    scanpaths_b200.models.loss.{LogAction, LogDuration, CrossEntropyLoss, MLPLogNormalDistribution}
    scanpaths_b200.models.baseline_attention.baseline
    scanpaths_b200.utils.evaltools.scanmatch.ScanMatch
    scanpaths_b200.utils.evaltools.visual_attention_metrics.{string_edit_distance, scaled_time_delay_embedding_similarity}
    scanpaths_b200.utils.evaluation.{evaluation, human_evaluation, pairs_eval, pairs_eval_scanmatch}
All of them call CUDA through the C ABI in include/scanpaths_b200.h; there is no
CPU fallback.
"""
__version__ = "0.1.0"
