"""osr.SpatialReference as used by dataset/utils.py:64-66 (a projection string the gdal shim ignores)."""


class SpatialReference:
    def ImportFromEPSG(self, code):
        self._code = code
        return 0

    def ExportToWkt(self):
        return f'GEOGCS["EPSG:{getattr(self, "_code", 4326)}"]'
