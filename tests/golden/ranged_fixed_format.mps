* fixed-format card layout (fields at columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61), names with blanks,
* every row type, RANGES on E / G / L rows, an objective constant, every bound type, a second N row.
NAME          RANGED
ROWS
 N  OBJ
 N  FREE ROW
 E  EQ 1
 G  GE 1
 L  LE 1
 E  EQ R
COLUMNS
    X 1       OBJ                1.0   EQ 1               1.0
    X 1       GE 1               2.0   FREE ROW           9.0
    X 2       OBJ               -2.0   EQ 1               1.0
    X 2       LE 1               3.0   EQ R               1.0
    X 3       GE 1               1.0   LE 1              -1.0
    X 4       EQ R               1.0
    X 5       OBJ                0.5
    X 6       LE 1               1.0
RHS
    RHS       OBJ               -7.5   EQ 1               4.0
    RHS       GE 1               1.0   LE 1              10.0
    RHS       EQ R               2.0
RANGES
    RNG       GE 1               2.5   LE 1               4.0
    RNG       EQ R              -3.0
BOUNDS
 UP BND       X 1                5.0
 MI BND       X 2
 FX BND       X 3                1.5
 FR BND       X 4
 UP BND       X 5               -1.0
 BV BND       X 6
ENDATA
