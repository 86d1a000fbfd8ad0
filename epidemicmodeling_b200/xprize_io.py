"""OxCGRT / XPRIZE wire formats on the host side of the path.

Readers for the files Tools/TrainPredictPrescribeNPI.m:62-90 loads with readtable -- the Oxford
time-series file (`CountryName, RegionName, Date (yyyymmdd), ConfirmedCases, ConfirmedDeaths,
<NPI columns>`), the populations file and the NPI cost file -- and a writer for the prescription
file of `xprize-sample-data/*_prescriptions_example.csv`
(`PrescriptionIndex, CountryName, RegionName, Date, <12 NPI columns>`).
Plain pandas: this is file plumbing, the arithmetic happens on the device (Engine.preprocess).
"""
import numpy as np
import pandas as pd

NPI_COLUMNS = ["C1_School closing", "C2_Workplace closing", "C3_Cancel public events",
               "C4_Restrictions on gatherings", "C5_Close public transport", "C6_Stay at home requirements",
               "C7_Restrictions on internal movement", "C8_International travel controls",
               "H1_Public information campaigns", "H2_Testing policy", "H3_Contact tracing",
               "H6_Facial Coverings"]                       # testPrescribeXPRIZE02.m:22-36
NPI_MAXES = np.array([3, 3, 2, 4, 2, 3, 2, 4, 2, 3, 2, 4], dtype=np.float64)   # :38


def geo_id(country, region):
    """strcat(CountryName, " ", RegionName) with an empty region (TrainPredictPrescribeNPI.m:66,83)."""
    region = "" if (region is None or (isinstance(region, float) and np.isnan(region))) else str(region)
    return f"{country} {region}"


def _date_number(s):
    return int(str(s).replace("-", ""))                     # :33-37 "2020-03-15" -> 20200315


def read_oxcgrt(path, start_date, end_date, npi_columns=NPI_COLUMNS, geo_ids=None):
    """Rows of every region between the two dates (inclusive, :102,131).
    Returns (ids, dates, cc [T, B], deaths [T, B], ip [T, L, B]); regions whose date range is
    incomplete are dropped (the batch needs one T)."""
    df = pd.read_csv(path, dtype={"CountryName": str, "RegionName": str}, keep_default_na=True)
    df["RegionName"] = df["RegionName"].fillna("")
    lo, hi = _date_number(start_date), _date_number(end_date)
    num = df["Date"].astype(str).str.replace("-", "", regex=False).astype(int)
    df = df[(num >= lo) & (num <= hi)].assign(_num=num)
    df["_id"] = df["CountryName"] + " " + df["RegionName"]
    dates = np.sort(df["_num"].unique())
    ids, cc, dd, ip = [], [], [], []
    for gid, g in df.groupby("_id", sort=False):             # unique(AllGeoIDs, 'stable') :86
        if geo_ids is not None and gid not in geo_ids:
            continue
        g = g.sort_values("_num")
        if len(g) != len(dates) or not np.array_equal(g["_num"].to_numpy(), dates):
            continue
        ids.append(gid)
        register_pair(g["CountryName"].iloc[0], g["RegionName"].iloc[0])
        cc.append(g["ConfirmedCases"].to_numpy(dtype=np.float64))
        dd.append(g["ConfirmedDeaths"].to_numpy(dtype=np.float64) if "ConfirmedDeaths" in g else np.full(len(g), np.nan))
        ip.append(g[list(npi_columns)].to_numpy(dtype=np.float64))
    if not ids:
        raise ValueError("no region covers the whole date range")
    return (ids, dates, np.ascontiguousarray(np.stack(cc, axis=1)), np.ascontiguousarray(np.stack(dd, axis=1)),
            np.ascontiguousarray(np.stack(ip, axis=2)))


def read_populations(path):
    df = pd.read_csv(path, dtype={"CountryName": str, "RegionName": str})
    df["RegionName"] = df["RegionName"].fillna("")
    return {c + " " + r: float(p) for c, r, p in zip(df["CountryName"], df["RegionName"], df["Population2020"])}


def read_costs(path, npi_columns=NPI_COLUMNS):
    df = pd.read_csv(path, dtype={"CountryName": str, "RegionName": str})
    df["RegionName"] = df["RegionName"].fillna("")
    return {c + " " + r: row for c, r, row in zip(df["CountryName"], df["RegionName"],
                                                   df[list(npi_columns)].to_numpy(dtype=np.float64))}


def write_prescriptions(path, ids, dates, schedules, npi_columns=NPI_COLUMNS):
    """schedules[index] = array [B, K, L] of NPI levels; dates = K 'yyyy-mm-dd' strings."""
    rows = []
    for idx, sch in enumerate(schedules):
        sch = np.asarray(sch)
        for b, gid in enumerate(ids):
            for k, d in enumerate(dates):
                rows.append([idx, _country_of(gid), _region_of(gid), d] + [int(round(v)) for v in sch[b, k]])
    pd.DataFrame(rows, columns=["PrescriptionIndex", "CountryName", "RegionName", "Date"] + list(npi_columns)) \
        .to_csv(path, index=False)


_SPLIT = {}


def register_pair(country, region):
    """Remember how an id splits back into (CountryName, RegionName) -- country names contain blanks."""
    _SPLIT[geo_id(country, region)] = (country, "" if region is None else region)


def _country_of(gid):
    return _SPLIT.get(gid, (gid.rstrip(), ""))[0]


def _region_of(gid):
    return _SPLIT.get(gid, (gid.rstrip(), ""))[1]
