from ._private.gcs_utils import GcsClient  # noqa: F401
