"""Drop-in for the reference's sert/models.py: the class surface lives in sert_b200.models."""
from sert_b200.models import (  # noqa: F401
    ModelInterface, ModelBase, LanguageModelBase, LanguageModel, VectorSpaceLanguageModelBase,
    VectorSpaceLanguageModel, LogLinearPredictFn, VectorSpacePredictFn, inproduct_sigmoid_distance,
    l2_regularization, glorot_uniform)
