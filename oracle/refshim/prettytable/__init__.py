class PrettyTable:
    def __init__(self, field_names=None):
        self.field_names = field_names or []
        self.rows = []
        self.align = {}
        self.sortby = None

    def add_row(self, row):
        self.rows.append(row)

    def __str__(self):
        return "\n".join(str(r) for r in [self.field_names] + self.rows)
