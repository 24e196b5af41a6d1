"""Name-compatible alias of the reference's probes/ package (probe.py, constructLabel.py, probe_trainModel.py, loss.py)."""
from .modules import OneWordPSDProbe, TwoWordPSDProbe, ConstructLabelGaget, probe, CombineLoss  # noqa: F401
