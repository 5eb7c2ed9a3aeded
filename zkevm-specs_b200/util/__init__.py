from .arithmetic import *  # noqa: F401,F403
from .arithmetic import FQ, RLC, Word, WordOrValue, Expression, linear_combine_bytes, add_words, FR_MODULUS  # noqa: F401
from .hash import keccak256, EMPTY_HASH, EMPTY_CODE_HASH, EMPTY_TRIE_HASH  # noqa: F401
