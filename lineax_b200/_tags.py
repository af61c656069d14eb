"""Operator tags (static metadata), mirroring lineax/_tags.py:24-87."""


class _HasRepr:
    def __init__(self, string: str):
        self.string = string

    def __repr__(self):
        return self.string


symmetric_tag = _HasRepr("symmetric_tag")
diagonal_tag = _HasRepr("diagonal_tag")
tridiagonal_tag = _HasRepr("tridiagonal_tag")
unit_diagonal_tag = _HasRepr("unit_diagonal_tag")
lower_triangular_tag = _HasRepr("lower_triangular_tag")
upper_triangular_tag = _HasRepr("upper_triangular_tag")
positive_semidefinite_tag = _HasRepr("positive_semidefinite_tag")
negative_semidefinite_tag = _HasRepr("negative_semidefinite_tag")

_SELF_TRANSPOSE = (
    symmetric_tag, unit_diagonal_tag, diagonal_tag, positive_semidefinite_tag,
    negative_semidefinite_tag, tridiagonal_tag,
)


def transpose_tags(tags: frozenset) -> frozenset:
    """Tags of the transposed operator (lineax/_tags.py:62-87)."""
    if symmetric_tag in tags:
        return tags
    out = [t for t in _SELF_TRANSPOSE if t in tags]
    if lower_triangular_tag in tags:
        out.append(upper_triangular_tag)
    if upper_triangular_tag in tags:
        out.append(lower_triangular_tag)
    return frozenset(out)
