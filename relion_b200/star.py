"""STAR metadata tables: the on-disk format of particle sets, optics groups and model files.

Host-side mirror of MetaDataTable's reader / writer (/root/reference/src/metadata_table.cpp):
  readStar      :1203-1267   "data_<name>" blocks; "# version N" tags; a block is a loop ("loop_") or a list of "_label value" pairs
  readStarLoop  :1036-1130   "_rlnLabel #n" header lines (the "#n" is a comment), then one row per line until an empty line;
                             more values than labels is an error, fewer too (except two-column tables)
  readStarList  :1132-1201   label / value pairs until "loop_", the next "data_" or the end of the file
  nextTokenInSTAR src/strings.cpp:595-660   blank-separated tokens, single- or double-quoted strings, "#" starts a comment
                             only at the start of a token
  write         :1366-1519   "\\n# version 50001\\n\\ndata_name\\n\\nloop_ \\n_rlnLabel #1 \\n...", values right-aligned to width 10
CR+LF files are refused like the reference does (:1223-1228).
"""
from __future__ import annotations

import io
import os
import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

CURRENT_VERSION = 50001

_INT_RE = re.compile(r"^[+-]?\d+$")
_FLOAT_RE = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$|^[+-]?(nan|inf)$", re.IGNORECASE)


class StarError(ValueError):
    pass


def tokens(line: str) -> List[str]:
    """Tokens of one STAR line (nextTokenInSTAR): blanks separate, quotes group, '#' at a token start ends the line."""
    out: List[str] = []
    i, n = 0, len(line)
    while i < n:
        while i < n and line[i] in " \t\n":
            i += 1
        if i >= n or line[i] == "#":
            break
        if line[i] in "'\"":
            q = line[i]
            i += 1
            j, val, closed = i, [], False
            while j < n:
                if line[j] == q and (j + 1 == n or line[j + 1] in " \t\n"):
                    closed = True
                    break
                val.append(line[j])
                j += 1
            if not closed:
                raise StarError(f"unterminated quoted string in STAR line: {line!r}")
            out.append("".join(val))
            i = j + 1
        else:
            j = i
            while j < n and line[j] not in " \t\n":
                j += 1
            out.append(line[i:j])
            i = j
    return out


@dataclass
class StarTable:
    """One data block.  `columns` keeps label order; a list block (is_list) has exactly one row."""
    name: str = ""
    columns: Dict[str, list] = field(default_factory=dict)
    is_list: bool = False
    version: int = 30000

    def __len__(self) -> int:
        return len(next(iter(self.columns.values()))) if self.columns else 0

    def labels(self) -> List[str]:
        return list(self.columns)

    def has(self, label: str) -> bool:
        return label in self.columns

    def column(self, label: str, dtype=None, default=None) -> np.ndarray:
        """Column as an array; `default` (scalar) is used when the label is absent (getValue returning false)."""
        if label not in self.columns:
            if default is None:
                raise KeyError(f"STAR table data_{self.name} has no column _{label}")
            return np.full(len(self), default, dtype=dtype or type(default))
        col = self.columns[label]
        return np.asarray(col, dtype=dtype) if dtype is not None else np.asarray(col)

    def value(self, label: str, default=None):
        """Single value of a list block (or of row 0)."""
        if label not in self.columns:
            if default is None:
                raise KeyError(f"STAR table data_{self.name} has no value _{label}")
            return default
        return self.columns[label][0]

    def set_column(self, label: str, values: Iterable) -> None:
        vals = list(values.tolist() if isinstance(values, np.ndarray) else values)
        if self.columns and len(vals) != len(self):
            raise StarError(f"column _{label}: {len(vals)} values for a table of {len(self)} rows")
        self.columns[label] = vals

    def sorted_by(self, label: str) -> "StarTable":
        """Stable sort on a string column (MetaDataTable::newSort, used on rlnMicrographName by Experiment::read)."""
        order = sorted(range(len(self)), key=lambda i: self.columns[label][i])
        return StarTable(self.name, {k: [v[i] for i in order] for k, v in self.columns.items()}, self.is_list, self.version)


def _convert(tok: str):
    if _INT_RE.match(tok):
        return int(tok)
    if _FLOAT_RE.match(tok):
        return float(tok)
    return tok


def _unify(col: list) -> list:
    """A column is int only if every entry parsed as int, float if all are numbers, else strings as written."""
    kinds = {type(v) for v in col}
    if kinds <= {int}:
        return col
    if kinds <= {int, float}:
        return [float(v) for v in col]
    return [v if isinstance(v, str) else repr(v) if isinstance(v, float) else str(v) for v in col]


def parse_star(text: str) -> List[StarTable]:
    """All data blocks of a STAR file, in file order (MetaDataTable::readAll)."""
    if "\r\n" in text:
        raise StarError("CR+LF line ends are not supported in STAR files (convert with dos2unix)")
    lines = text.split("\n")
    tables: List[StarTable] = []
    version = 30000
    i, n = 0, len(lines)
    while i < n:
        line = lines[i].strip()
        i += 1
        if "# version " in line:
            try:
                version = int(line.split("# version ", 1)[1].split()[0])
            except (ValueError, IndexError):
                pass
            continue
        if not line.startswith("data_"):
            continue
        t = StarTable(name=line[5:].strip(), version=version)
        tables.append(t)
        raw_cols: Dict[str, list] = {}
        # list part: "_label value" pairs
        while i < n:
            cur = lines[i].strip()
            if cur.startswith("loop_") or cur.startswith("data_"):
                break
            i += 1
            if not cur or cur[0] in "#;":
                if "# version " in cur:
                    try:
                        version = int(cur.split("# version ", 1)[1].split()[0])
                    except (ValueError, IndexError):
                        pass
                continue
            if cur[0] == "_":
                tk = tokens(cur)
                if len(tk) < 2:
                    raise StarError(f"STAR list entry without a value: {cur!r}")
                raw_cols[tk[0][1:]] = [tk[1]]
                t.is_list = True
        if i < n and lines[i].strip().startswith("loop_"):
            i += 1
            labels: List[str] = []
            while i < n:
                cur = lines[i].strip()
                if not cur or cur[0] in "#;":
                    i += 1
                    continue
                if cur[0] != "_":
                    break
                labels.append(cur[1:].split("#")[0].split()[0])
                i += 1
            cols = [[] for _ in labels]
            while i < n:
                cur = lines[i].strip()
                if not cur:
                    break
                if cur.startswith("data_"):
                    break
                i += 1
                if cur[0] == "#":
                    continue
                tk = tokens(cur)
                if len(tk) > len(labels):
                    raise StarError("A line in the STAR file contains more columns than the number of labels: " + cur)
                if len(tk) < len(labels):
                    if len(labels) > 2:
                        raise StarError(f"A line in the STAR file contains fewer columns than the number of labels. "
                                        f"Expected = {len(labels)} Found = {len(tk)}: {cur}")
                    tk = tk + [""] * (len(labels) - len(tk))
                for c, v in zip(cols, tk):
                    c.append(v)
            if labels:
                t.is_list = False
                raw_cols = {lab: c for lab, c in zip(labels, cols)}
        t.columns = {k: _unify([_convert(v) for v in col]) for k, col in raw_cols.items()}
    return tables


def read_star(path: str, name: Optional[str] = None):
    """All tables of a file as {name: table}, or the one called `name` (the first block when name == "")."""
    with open(path, "r") as f:
        tables = parse_star(f.read())
    if name is None:
        return {t.name: t for t in tables}
    for t in tables:
        if name == "" or t.name == name:
            return t
    raise KeyError(f"{path} has no data_{name} block")


def _escape(v) -> str:
    """Value as MetaDataTable::getValueToString prints it (:228-290): doubles %12.6f, scientific outside [1e-3, 1e5], one
    digit less when negative; integers %12ld; strings quoted when they need it."""
    if isinstance(v, (bool, np.bool_)):
        return "%12d" % int(v)
    if isinstance(v, (int, np.integer)):
        return "%12d" % int(v)
    if isinstance(v, (float, np.floating)):
        v = float(v)
        sci = (0.0 < abs(v) < 0.001) or abs(v) > 100000.0
        fmt = ("%12.5e" if v < 0 else "%12.6e") if sci else ("%12.5f" if v < 0 else "%12.6f")
        return (fmt % v)[:12]                             # snprintf(buffer, 13, ...) truncates to 12 characters
    s = str(v)
    if s == "" or any(ch in s for ch in " \t") or s[0] in "'\"#_;":
        q = '"' if '"' not in s else "'"
        return q + s + q
    return s


def format_star(tables: Sequence[StarTable]) -> str:
    out = io.StringIO()
    for t in tables:
        if not t.columns:
            continue                                      # "Only write tables that have something in them"
        out.write(f"\n# version {CURRENT_VERSION}\n\ndata_{t.name}\n\n")
        if t.is_list:
            for lab, col in t.columns.items():
                out.write(f"_{lab:<40s} {_escape(col[0]):>12s}\n")
            out.write(" \n")
        else:
            out.write("loop_ \n")
            for k, lab in enumerate(t.columns, start=1):
                out.write(f"_{lab} #{k} \n")
            cols = list(t.columns.values())
            for r in range(len(t)):
                out.write(" ".join(f"{_escape(c[r]):>10s}" for c in cols) + " \n")
            out.write(" \n")
    return out.getvalue()


def write_star(path: str, tables: Sequence[StarTable]) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        f.write(format_star(tables))
    os.replace(tmp, path)
