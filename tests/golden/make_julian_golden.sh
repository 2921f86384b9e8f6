#!/bin/bash
# Generates tests/golden/julian_dates.sha256 from the UNMODIFIED reference: tests/cpp/julian_dates.cpp
# built against the reference's own headers and sources (needs /root/reference; run from the repo root).
set -e
REF=${REF:-/root/reference}
g++ -std=c++11 -w -DPERTURB_BATCH_USE_PERTURB_HPP -I$REF/include -Iinclude tests/cpp/julian_dates.cpp \
    $REF/src/perturb.cpp $REF/src/sgp4.cpp $REF/src/tle.cpp -o /tmp/julian_dates_ref
/tmp/julian_dates_ref | sha256sum | cut -d' ' -f1 > tests/golden/julian_dates.sha256
/tmp/julian_dates_ref | wc -l >> tests/golden/julian_dates.sha256
cat tests/golden/julian_dates.sha256
