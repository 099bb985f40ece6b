#!/bin/bash
# The reference's run.sh with the B200 executable in place of bin/speedy: run it from the root of a speedy.f90 checkout
# (`bash <repo>/tools/run_b200.sh`).  Same steps: a fresh run directory, the seven boundary files linked in, namelist.nml copied,
# the model run with its output kept in output.txt.  The executable reads the NetCDF-4 boundary files where they are linked.
#   RUNDIR=/somewhere/else bash tools/run_b200.sh      # when the checkout is read-only
# Outside a checkout (no data/bc/t30 here) the packed copy of this repository is used, and the reference's defaults apply when
# there is no namelist.nml.  Extra arguments go to the executable (--members 8 --sppt --member -1, --trunc 47 ...).
REPO=$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)
EXE=$REPO/speedy.f90_b200/bin/speedy_b200
ROOT=`pwd`
RUNDIR=${RUNDIR:-$ROOT/rundir}
CLIM=$ROOT/data/bc/t30/clim
ANOM=$ROOT/data/bc/t30/anom

if [ ! -x $EXE ]; then
    echo "No executable found ($EXE)"
    echo "Have you run make -C $REPO/speedy.f90_b200 yet?"
    exit 1
fi

# Make and move to run directory
rm -rf $RUNDIR
mkdir -p $RUNDIR
cd $RUNDIR || exit 1

BC=()
if [ -f $CLIM/surface.nc ]; then
    # Link input files
    for f in surface sea_surface_temperature sea_ice land snow soil; do ln -s $CLIM/$f.nc .; done
    ln -s $ANOM/sea_surface_temperature_anomaly.nc .
else
    BC=(--bc $REPO/data/bc_t30.bin)
fi

# Copy namelist file to run directory
[ -f $ROOT/namelist.nml ] && cp $ROOT/namelist.nml $RUNDIR

# Run SPEEDY
time $EXE "${BC[@]}" "$@" | tee output.txt
exit ${PIPESTATUS[0]}
