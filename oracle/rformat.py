"""R number formatting / key building, restated (oracle; test infrastructure).

Follows:
  * ``signif(x)`` (6 digits)  -- R nmath/fprec.c, used at
    /root/reference/R/training.R:690-691, prediction.R:537,939.
  * ``tidyr::unite(sep=":")`` -> ``as.character(double)``: 15 significant
    digits, fixed notation unless scientific is strictly shorter
    (R format.c scientific()/formatReal()).
  * ``dplyr::arrange`` = C-locale (byte-wise) string order
    (/root/reference/R/training.R:699-701, em-magma.R:27-30).
"""
import math

import numpy as np


def r_signif(x, digits=6):
    """R's fprec(): round to `digits` significant digits."""
    x = float(x)
    if not math.isfinite(x) or x == 0.0:
        return x
    dig = int(round(digits))
    if dig > 22:
        return x
    if dig < 1:
        dig = 1
    sgn = 1.0
    if x < 0.0:
        sgn = -1.0
        x = -x
    l10 = math.log10(x)
    e10 = int(dig - 1 - math.floor(l10))
    p10 = 1.0
    if e10 > 308:                       # max10e: split the power so it stays finite
        p10 = 10.0 ** (e10 - 308)
        e10 = 308
    if e10 > 0:
        pow10 = 10.0 ** e10
        return sgn * (float(np.rint((x * pow10) * p10)) / pow10) / p10
    pow10 = 10.0 ** (-e10)
    return sgn * (float(np.rint(x / pow10)) * pow10)


def r_as_character(x):
    """as.character(<double>) with R's 15-significant-digit rule."""
    x = float(x)
    if math.isnan(x):
        return "NA"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "0"
    neg = x < 0
    ref = float("%.14e" % x)
    nsig = 15
    for n in range(1, 16):
        if float("%.*e" % (n - 1, x)) == ref:
            nsig = n
            break
    mant, exp = ("%.*e" % (nsig - 1, abs(x))).split("e")
    e = int(exp)
    digits = mant.replace(".", "")
    # widths as in formatReal(): fixed is used when not wider than scientific
    wexp = 2 if abs(e) < 100 else 3
    w_sci = neg + (nsig + 1 if nsig > 1 else 1) + 2 + wexp
    left = e + 1 if e >= 0 else 1
    rgt = max(0, nsig - e - 1)
    w_fix = neg + left + (rgt + 1 if rgt > 0 else 0)
    if w_fix <= w_sci:
        if e >= 0:
            intpart = digits[: e + 1].ljust(e + 1, "0")
            frac = digits[e + 1:]
        else:
            intpart = "0"
            frac = "0" * (-e - 1) + digits
        s = intpart + ("." + frac if rgt > 0 else "")
    else:
        s = mant + "e" + ("-" if e < 0 else "+") + ("%0*d" % (wexp, abs(e)))
    return ("-" if neg else "") + s


def reference_key(row):
    """paste(signif(cols), collapse=":") for one input row (iterable)."""
    return ":".join(r_as_character(r_signif(v)) for v in row)


def c_locale_order(keys):
    """Indices sorting `keys` byte-wise (dplyr::arrange C locale), stable."""
    return sorted(range(len(keys)), key=lambda i: keys[i].encode("utf-8"))
