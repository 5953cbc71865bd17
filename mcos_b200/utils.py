"""Price-history adaptor (mcos/utils.py:6-20). Host-side, runs once before the loop: annualised mean daily returns and
pairwise-complete sample covariance, as PyPortfolioOpt 0.5.2's mean_historical_return / sample_cov compute them."""
import pandas as pd


def convert_price_history(df: pd.DataFrame, frequency: int = 252):
    daily = df.pct_change(fill_method=None).dropna(how="all")
    mu = daily.mean() * frequency
    cov = daily.cov() * frequency
    return mu.to_numpy(), cov.to_numpy()
