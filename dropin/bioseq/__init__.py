"""Drop-in shim for ``import bioseq``: re-exports the tokenizer surface of bioseq_b200
(the reference's model zoo, loaders and POA utilities are out of scope)."""
import cbioseq  # noqa: F401
from cbioseq import *  # noqa: F401,F403
from bioseq_b200 import *  # noqa: F401,F403
from bioseq_b200 import __all__  # noqa: F401
