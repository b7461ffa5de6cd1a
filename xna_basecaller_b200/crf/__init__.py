"""Plugin entry point: `[model] package = "xna_basecaller_b200.crf"` in a model's config.toml makes the reference's
load_symbol (ub-bonito/bonito/util.py:228-239) pick up Model and basecall from here (bonito/crf/__init__.py)."""
from .model import Model
from .basecall import basecall
