"""Drop-in for this variant's custom_components/custom_callbacks.py; see custom_components/_callbacks.py."""
from ...custom_components._callbacks import AccuracyLossTrackingCallBack, NestedScaleTrackingCallback  # noqa: F401
