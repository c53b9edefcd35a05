"""TEST-ONLY stand-in for influxdb-client (see tests/stubs/README.md).  dps.py creates a client at import time
(src/influx_handler.py:41-42) and only writes when the server answers a TCP probe, which it never does in the tests."""
WRITES = []


class Point:
    def __init__(self, name):
        self.name, self.tags, self.fields = name, {}, {}

    def tag(self, k, v):
        self.tags[k] = v
        return self

    def field(self, k, v):
        self.fields[k] = v
        return self


class _WriteApi:
    def write(self, bucket=None, record=None, **kwargs):
        WRITES.append((bucket, record))


class InfluxDBClient:
    def __init__(self, url=None, token=None, org=None, **kwargs):
        self.url, self.token, self.org = url, token, org

    def write_api(self, *args, **kwargs):
        return _WriteApi()

    def close(self):
        pass
