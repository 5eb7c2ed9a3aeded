"""Import-only stub: secp256k1 ECDSA is third-party curve math, out of scope."""


class KeyAPI:
    class PublicKey:
        def __init__(self, *a, **k):
            raise NotImplementedError("eth_keys is stubbed in the oracle shim")

    class PrivateKey(PublicKey):
        pass

    class Signature(PublicKey):
        pass

    def __init__(self, *a, **k):
        pass


keys = KeyAPI()
