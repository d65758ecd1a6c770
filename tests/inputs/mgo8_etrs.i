set cell 7.70 0 0 0 7.70 0 0 0 7.70
species magnesium Mg.xml
species oxygen O.xml
atom Mg1 magnesium 0.00 0.00 0.00
atom Mg2 magnesium 0.00 3.85 3.85
atom Mg3 magnesium 3.85 0.00 3.85
atom Mg4 magnesium 3.85 3.85 0.00
atom O1 oxygen 3.85 0.00 0.00
atom O2 oxygen 0.00 3.85 0.00
atom O3 oxygen 0.00 0.00 3.85
atom O4 oxygen 3.85 3.85 3.85
set force_complex_wf ON
set ecut 50
set xc LDA
set wf_dyn PSDA
set ecutprec 8
randomize_wf
run 0 30
set wf_dyn ETRS
set TD_dt 0.05
run 5 1
quit
