class RegularPolygon:  # pragma: no cover
    pass
