// The reference's generated objective for the introduction-single-stock model, compiled where it lies.
#include "/root/reference/baseline/introduction-single-stock.cpp"
#include "driver.inc"
