class _Cell:
    def __init__(self):
        self.value = None
        self.number_format = None


class _AutoFilter:
    ref = None


class Worksheet:
    def __init__(self, title="Sheet"):
        self.title = title
        self.cells = {}
        self.freeze_panes = None
        self.auto_filter = _AutoFilter()

    def cell(self, row, column):
        return self.cells.setdefault((row, column), _Cell())

    def merge_cells(self, **kwargs):
        pass


class Workbook:
    def __init__(self, **kwargs):
        self.worksheets = [Worksheet()]

    @property
    def sheetnames(self):
        return [w.title for w in self.worksheets]

    def __getitem__(self, name):
        return next(w for w in self.worksheets if w.title == name)

    def __delitem__(self, name):
        self.worksheets = [w for w in self.worksheets if w.title != name]

    def index(self, ws):
        return self.worksheets.index(ws)

    def remove(self, ws):
        self.worksheets.remove(ws)

    def create_sheet(self, title=None, index=None):
        ws = Worksheet(title or "Sheet%d" % (len(self.worksheets) + 1))
        self.worksheets.insert(len(self.worksheets) if index is None else index, ws)
        return ws

    def save(self, handle):
        lines = []
        for ws in self.worksheets:
            lines.append("# sheet %s" % ws.title)
            rows = sorted({r for r, _ in ws.cells})
            cols = sorted({c for _, c in ws.cells})
            for r in rows:
                lines.append("\t".join("" if (r, c) not in ws.cells or ws.cells[(r, c)].value is None else str(ws.cells[(r, c)].value) for c in cols))
        data = ("\n".join(lines) + "\n").encode("utf-8")
        if hasattr(handle, "write"):
            handle.write(data)
        else:
            with open(handle, "wb") as f:
                f.write(data)
