"""TEST-ONLY stand-in for flask (see tests/stubs/README.md): dps.py imports src/api_handler.py, which builds an app
and registers two routes at import time; the tests answer 'n' to "start the REST API?", so nothing is ever served."""


class _Request:
    environ = {}


request = _Request()


def jsonify(*args, **kwargs):
    return args[0] if args else kwargs


class Flask:
    def __init__(self, name, **kwargs):
        self.name, self.routes = name, {}

    def route(self, rule, **options):
        def deco(fn):
            self.routes[rule] = fn
            return fn
        return deco

    def response_class(self, body, content_type=None):
        return body

    def run(self, **kwargs):
        raise RuntimeError("test stub: the REST API is not served in tests")
