#Population_format_version 0.0.1
id size contactDensity conDenAfterLD,startLD,endLD samplingMultiplier
0 200000 1.0 0.2,0.02,0.004 1.5
1 100000 0.9 2.0 0.3,0.03,0.005
2 50000 1.1 0.5
3 40000 1.2 0.4,0.05,0.01
