"""B200-native (sm_100a) implementation of ub-bonito's basecalling forward path.

Drop-in plugin surface (select with ``[model] package = "xna_basecaller_b200.crf"`` in a model's
config.toml, the hook at bonito/util.py:228-239):

    xna_basecaller_b200.nn           Serial / Convolution / LSTM / LinearCRFEncoder / Permute / Reverse / Swish
    xna_basecaller_b200.crf          Model, basecall            (bonito/crf/__init__.py)
    xna_basecaller_b200.crf.model    CTC_CRF, SeqdistModel, Model
    xna_basecaller_b200.crf.basecall compute_scores, stitch_results, basecall
    xna_basecaller_b200.util         chunk, stitch, batchify, unbatchify, load_symbol, load_model

Every compute stage is a hand-written CUDA kernel in libxna_b200.so behind the C ABI of
include/xna_basecaller.h; there is no CPU or PyTorch fallback.
"""
__version__ = '0.1.0'
