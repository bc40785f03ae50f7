from dbsgym_b200.utils import *  # noqa: F401,F403
from dbsgym_b200.utils import generate_w0_with_locus  # noqa: F401
