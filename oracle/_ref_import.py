"""TEST INFRASTRUCTURE ONLY. Imports the unmodified reference from /root/reference (authoring
container only; the path does not exist on the GPU box) with stub modules for the four packages the
reference imports but this image lacks: boto3, webdataset, albumentations, braceexpand.
Used solely by oracle/gen_golden.py to pin the oracle restatement against the live reference."""
import sys, types
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


def import_reference():
    for name in ["boto3", "boto3.s3", "boto3.s3.transfer", "webdataset", "webdataset.handlers",
                 "webdataset.filters", "albumentations", "braceexpand"]:
        if name not in sys.modules:
            m = _Stub(name)
            m.__path__ = []
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import egom2p.models.egom2p_model as ref_model  # noqa
    return ref_model
