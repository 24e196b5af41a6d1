"""mtvaf_b200 -- B200-native (sm_100a) implementation of the MTVAF data-parallel hot path.

Importing the package loads the CUDA extension (mtvaf_b200/_C/libmtvaf_b200.so) and fails loudly if it
is missing: there is no CPU or PyTorch fallback.  `mtvaf_b200.synthetic` (pure torch-CPU data
generation) can be imported on its own without the extension.
"""
__all__ = ["RobertaModel", "BertModel", "TVNetSAModel", "TVNetSAModel2", "CRF", "probe", "OneWordPSDProbe", "TwoWordPSDProbe",
           "ConstructLabelGaget", "CombineLoss", "FeatureStub", "ImageModel"]


def __getattr__(name):
    if name in __all__:
        from . import modules
        return getattr(modules, name)
    raise AttributeError(name)
