"""Sample / pair ids exactly as the reference derives them (Guava Hashing.md5: little-endian primitive sinks,
asLong = first 8 digest bytes little-endian, signed)."""
import hashlib
import struct


def pair_id(left_id: int, right_id: int) -> int:
    """md5.putLong(left).putLong(right).asLong — GenometricMap71.scala:174, GenometricJoin.scala:46, CombineMD.scala:68."""
    return int.from_bytes(hashlib.md5(struct.pack("<qq", left_id, right_id)).digest()[:8], "little", signed=True)


def long_id(v: int) -> int:
    """md5.putLong(v).asLong — GenometricDifference.scala:53 (output sample id of DIFFERENCE)."""
    return int.from_bytes(hashlib.md5(struct.pack("<q", v)).digest()[:8], "little", signed=True)


def string_id(s: str) -> int:
    return int.from_bytes(hashlib.md5(s.encode("utf-8")).digest()[:8], "little", signed=True)


def sample_id_from_filename(name: str) -> int:
    """loaders/Loaders.scala:203-208."""
    ext = name[name.rfind(".") + 1:]
    base = name if ext != "meta" else name[: name.rfind(".")]
    return string_id(base.replace("/", ""))
